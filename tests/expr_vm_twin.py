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


def valid(instructions, n_regs, n_params, n_out):
    for op, dst, a, b in np.asarray(instructions).tolist():
        if op == mu.OP_OUT:
            ok = 0 <= dst < n_out and 0 <= a < n_regs
        elif op == mu.OP_CONST:
            ok = 0 <= dst < n_regs
        elif op == mu.OP_PARAM:
            ok = 0 <= dst < n_regs and 0 <= a < n_params
        elif op in _UNARY or op == mu.OP_POWI:
            ok = 0 <= dst < n_regs and 0 <= a < n_regs
        elif op in _BINARY:
            ok = 0 <= dst < n_regs and 0 <= a < n_regs and 0 <= b < n_regs
        else:
            ok = False
        if not ok:
            return False
    return True


def run(instructions, n_regs, mat_sizes, params):
    """params [B, P] -> flat buffer laid out like hmpc_param_eval_f64's out (matrix m = [B, size_m] block)."""
    params = np.asarray(params, dtype=np.float64)
    B, P = params.shape
    n_out = int(sum(mat_sizes))
    slots = np.full((n_out, B), np.nan)
    if valid(instructions, n_regs, P, n_out):
        regs = np.zeros((n_regs, B))
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
    out, off = np.empty(B * n_out), 0
    for sz in mat_sizes:
        out[B * off:B * (off + sz)] = slots[off:off + sz].T.reshape(-1)
        off += sz
    return out


def run_program(prog, params):
    """ExprProgram + params [B, P] -> dict name -> [B, rows, cols]."""
    flat = run(prog.instructions, prog.n_regs, prog.mat_sizes, params)
    B = np.asarray(params).shape[0]
    out, off = {}, 0
    for name, (r, c), sz in zip(prog.mat_names, prog.mat_shapes, prog.mat_sizes):
        out[name] = flat[B * off:B * (off + sz)].reshape(B, r, c)
        off += sz
    return out
