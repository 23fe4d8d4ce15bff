"""TEST INFRASTRUCTURE: numpy twin of the instruction interpreter in csrc/param_eval.cu (same opcodes, same
operand conventions, same validation).  Lets the CPU suite check the sympy -> register-program compiler
(pyhybridcontrol_b200/utils/matrix_utils.py) against the reference's lambdify results without a GPU; the GPU tests
then check the kernel against the same fixtures and against this twin on random programs."""
import struct

import numpy as np

from pyhybridcontrol_b200.utils import matrix_utils as mu

_UNARY = {mu.OP_MOV: lambda x: x, mu.OP_NEG: np.negative, mu.OP_ABS: np.abs, mu.OP_SIGN: np.sign, mu.OP_SQRT: np.sqrt,
          mu.OP_EXP: np.exp, mu.OP_LOG: np.log, mu.OP_SIN: np.sin, mu.OP_COS: np.cos, mu.OP_TAN: np.tan,
          mu.OP_ASIN: np.arcsin, mu.OP_ACOS: np.arccos, mu.OP_ATAN: np.arctan, mu.OP_SINH: np.sinh,
          mu.OP_COSH: np.cosh, mu.OP_TANH: np.tanh, mu.OP_FLOOR: np.floor, mu.OP_CEIL: np.ceil}
_BINARY = {mu.OP_ADD: np.add, mu.OP_SUB: np.subtract, mu.OP_MUL: np.multiply, mu.OP_DIV: np.divide,
           mu.OP_POW: np.power, mu.OP_MIN: np.minimum, mu.OP_MAX: np.maximum, mu.OP_ATAN2: np.arctan2}


def _powi(x, n):
    m, r, p = abs(int(n)), np.ones_like(x), x.copy()
    while m:
        if m & 1:
            r = r * p
        m >>= 1
        if m:
            p = p * p
    return 1.0 / r if n < 0 else r


def valid(instructions, n_regs, n_params, n_out, preload=False):
    """the checks the kernels make while loading a program; preload = the experimental kernel's form (parameters are
    registers 0..P-1, never written, no PARAM instruction)"""
    lo = n_params if preload else 0
    for op, dst, a, b in np.asarray(instructions).tolist():
        if op == mu.OP_OUT:
            ok = 0 <= dst < n_out and 0 <= a < n_regs
        elif op == mu.OP_CONST:
            ok = lo <= dst < n_regs
        elif op == mu.OP_PARAM:
            ok = (not preload) and 0 <= dst < n_regs and 0 <= a < n_params
        elif op in _UNARY or op == mu.OP_POWI:
            ok = lo <= dst < n_regs and 0 <= a < n_regs
        elif op in _BINARY:
            ok = lo <= dst < n_regs and 0 <= a < n_regs and 0 <= b < n_regs
        else:
            ok = False
        if not ok:
            return False
    return True


# instructions whose CUDA implementation is a library function (<= 2 ulp), not a correctly rounded IEEE operation
LIBRARY_OPS = frozenset((mu.OP_EXP, mu.OP_LOG, mu.OP_SIN, mu.OP_COS, mu.OP_TAN, mu.OP_ASIN, mu.OP_ACOS, mu.OP_ATAN,
                         mu.OP_SINH, mu.OP_COSH, mu.OP_TANH, mu.OP_POW, mu.OP_ATAN2))


def run(instructions, n_regs, mat_sizes, params, noise=None, noise_ulps=2, preload=False):
    """params [B, P] -> flat buffer laid out like hmpc_param_eval_f64's out (matrix m = [B, size_m] block).

    ``noise``: a numpy Generator -- every library-function result is moved by a random integer number of ulps in
    [-noise_ulps, noise_ulps] -- or an int: every such result is moved by that many ulps.  Several such runs give the envelope inside which two correct implementations of the
    same program may differ (a badly conditioned expression, e.g. acos(tanh(big)), amplifies one ulp a lot)."""
    params = np.asarray(params, dtype=np.float64)
    B, P = params.shape
    n_out = int(sum(mat_sizes))
    slots = np.full((n_out, B), np.nan)
    if valid(instructions, n_regs, P, n_out, preload=preload):
        regs = np.zeros((n_regs, B))
        if preload:
            regs[:P] = params.T
        with np.errstate(all="ignore"):
            for op, dst, a, b in np.asarray(instructions).tolist():
                if op == mu.OP_OUT:
                    slots[dst] = regs[a]
                elif op == mu.OP_CONST:
                    regs[dst] = struct.unpack("<d", struct.pack("<ii", a, b))[0]
                elif op == mu.OP_PARAM:
                    regs[dst] = params[:, a]
                elif op == mu.OP_POWI:
                    regs[dst] = _powi(regs[a], b)
                elif op in _UNARY:
                    regs[dst] = _UNARY[op](regs[a])
                else:
                    regs[dst] = _BINARY[op](regs[a], regs[b])
                if noise is not None and op in LIBRARY_OPS:
                    ulps = noise if isinstance(noise, int) else noise.integers(-noise_ulps, noise_ulps + 1, size=B)
                    regs[dst] = regs[dst] * (1.0 + ulps * 2.0 ** -52)
                    if op in (mu.OP_SIN, mu.OP_COS, mu.OP_TANH):          # no implementation leaves [-1, 1]
                        regs[dst] = np.clip(regs[dst], -1.0, 1.0)
    out, off = np.empty(B * n_out), 0
    for sz in mat_sizes:
        out[B * off:B * (off + sz)] = slots[off:off + sz].T.reshape(-1)
        off += sz
    return out


def run_program(prog, params, noise=None, noise_ulps=2):
    """ExprProgram + params [B, P] -> dict name -> [B, rows, cols]."""
    flat = run(prog.instructions, prog.n_regs, prog.mat_sizes, params, noise=noise, noise_ulps=noise_ulps)
    B = np.asarray(params).shape[0]
    out, off = {}, 0
    for name, (r, c), sz in zip(prog.mat_names, prog.mat_shapes, prog.mat_sizes):
        out[name] = flat[B * off:B * (off + sz)].reshape(B, r, c)
        off += sz
    return out


def envelope(prog, params, runs=8, seed=0, noise_ulps=2):
    """dict name -> [B, rows, cols]: largest deviation from the noise-free run over the coherent runs (every library
    result moved by -2, -1, +1, +2 ulp: catches every single sensitive operation) and ``runs`` random ones."""
    ref = run_program(prog, params)
    rng = np.random.default_rng(seed)
    env = {k: np.zeros_like(v) for k, v in ref.items()}
    modes = [u for u in range(-noise_ulps, noise_ulps + 1) if u] + [rng] * runs
    for mode in modes:
        out = run_program(prog, params, noise=mode, noise_ulps=noise_ulps)
        for k in env:
            with np.errstate(all="ignore"):
                dev = np.abs(out[k] - ref[k])
            env[k] = np.fmax(env[k], np.where(np.isfinite(dev), dev, np.inf))
    return ref, env
