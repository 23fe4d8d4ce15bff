"""Times hmpc_param_eval_f64 on the DEWH models (closed-form program of this package and the reference's own pinv()
expression from the golden fixture) at fleet sizes 1e2 .. 2e6, CUDA events on the launching stream, inputs resident.
Prints one JSON line per case: agents, program size, us per launch, algorithmic GB/s (8 (P + n_out) bytes per agent)
and its fraction of the measured HBM copy peak (MEASURED_PEAKS.json, else the profiling guide's fallback)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as M  # noqa: E402
from pyhybridcontrol_b200.utils.matrix_utils import ExprProgram  # noqa: E402


def peak_gbs():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for key in ("hbm_copy_gbs_burst", "hbm_gbs_burst", "hbm_copy_gbs", "hbm_gbs"):
            if key in pk:
                return float(pk[key]), "MEASURED_PEAKS.json:" + key
        for v in pk.values():
            if isinstance(v, dict):
                for key, x in v.items():
                    if "gb" in key.lower() and isinstance(x, (int, float)):
                        return float(x), "MEASURED_PEAKS.json:" + key
    except Exception:
        pass
    return 6458.7, "fallback"


def main():
    dev = torch.device("cuda:0")
    peak, src = peak_gbs()
    from test_callable_front_end import load_fixture
    z, names, mats, pnames = load_fixture("callable_dewh_sim.npz")
    progs = {"dewh_sim_closed_form": M.DewhModel.get_dewh_mld_symbolic(const_heat=False).to_callable().program,
             "dewh_control_closed_form": M.DewhModel.get_dewh_mld_symbolic(const_heat=True).to_callable().program,
             "dewh_sim_reference_pinv_expression": ExprProgram(mats, param_names=pnames)}
    rng = np.random.default_rng(0)
    for tag, prog in progs.items():
        base = np.array([float(M._par.dewh_param_struct[n]) for n in prog.param_names])
        for B in (100, 10000, 100000, 2000000):
            tab = np.tile(base, (B, 1)) * (1.0 + 0.05 * rng.uniform(-1, 1, size=(B, len(base))))
            if "T_h" in prog.param_names:
                tab[:, prog.param_names.index("T_h")] = rng.uniform(30, 80, B)
            params = torch.as_tensor(tab).to(dev)
            for _ in range(3):
                prog.evaluate(params)
            torch.cuda.synchronize()
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                prog.evaluate(params)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            gbs = prog.bytes_per_agent() * B / (us * 1e-6) / 1e9
            print(json.dumps(dict(kernel="param_eval_kernel", program=tag, agents=B, n_ins=prog.n_ins,
                                  n_regs=prog.n_regs, bytes_per_agent=prog.bytes_per_agent(), us_per_launch=round(us, 2),
                                  agents_per_s=round(B / (us * 1e-6)), achieved_gbs=round(gbs, 1), peak_gbs=peak,
                                  peak_source=src, frac=round(gbs / peak, 4))))


if __name__ == "__main__":
    main()
