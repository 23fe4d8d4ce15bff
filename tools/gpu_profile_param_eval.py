"""Three launches of hmpc_param_eval_f64 (DEWH simulation model, closed-form program, 2,000,000 agents) for ncu:
  ncu --set full --clock-control none -k regex:param_eval -c 1 -o gpurun_out/param_eval python tools/gpu_profile_param_eval.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as M  # noqa: E402

prog = M.DewhModel.get_dewh_mld_symbolic(const_heat=False).to_callable().program
B = 2000000
rng = np.random.default_rng(0)
base = np.array([float(M._par.dewh_param_struct[n]) for n in prog.param_names])
tab = base[None, :] * (1.0 + 0.05 * rng.uniform(-1, 1, size=(B, len(base))))
tab[:, prog.param_names.index("T_h")] = rng.uniform(30, 80, B)
params = torch.as_tensor(tab).to("cuda:0")
for _ in range(3):
    prog.evaluate(params)
torch.cuda.synchronize()
print("ok", prog.n_ins, prog.n_regs, prog.bytes_per_agent())
