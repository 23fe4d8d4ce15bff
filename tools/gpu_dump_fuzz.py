"""Runs the fuzz programs of tests/test_gpu_callable.py through hmpc_param_eval_f64 and saves the raw kernel outputs
(gpurun_out/fuzz_outputs.npz), so that the comparison against the numpy twin can be examined off the GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_callable as T  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    data = {}
    for trial, prog, params in T.fuzz_cases():
        out = T._run(prog, params, dev)
        for k, v in out.items():
            data["t%d_%s" % (trial, k)] = v
        bad, checked, tight = T.fuzz_check(trial, prog, params, out)
        print(trial, checked, tight, bad)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fuzz_outputs.npz"), **data)


if __name__ == "__main__":
    main()
