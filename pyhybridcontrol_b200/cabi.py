"""ctypes binding of ``csrc/libhmpc.so`` (C ABI declared in ``include/hmpc.h``).

This is the only door between the Python host code and the CUDA kernels.  PyTorch tensors are used purely as
device buffers (``tensor.data_ptr()``) and for the current stream.  There is NO CPU fallback: if the library
has not been built, importing this module raises ``ImportError``.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhmpc.so")

MAT_NAMES = ("A", "B1", "B2", "B3", "B4", "b5", "C", "D1", "D2", "D3", "D4", "d5",
             "E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")
EVO_NAMES = ("Phi_x", "Gamma_v", "Gamma_omega", "Gamma_5", "L_x", "L_v", "L_omega", "L_5",
             "H_x", "H_v", "H_omega", "H_5")
EXPORTS = ("hmpc_version", "hmpc_last_cuda_error", "hmpc_device_info", "hmpc_condense_f64",
           "hmpc_condense_bytes_per_agent", "hmpc_constraint_rhs_f64", "hmpc_predict_f64", "hmpc_linear_cost_f64",
           "hmpc_milp_default_opts", "hmpc_milp_workspace_bytes", "hmpc_milp_solve_f64", "hmpc_miqp_default_opts", "hmpc_miqp_workspace_bytes", "hmpc_miqp_solve_f64",
           "hmpc_stage_dp_default_opts",
           "hmpc_stage_dp_supported", "hmpc_stage_dp_workspace_bytes", "hmpc_stage_dp_max_cells", "hmpc_stage_dp_solve_f64", "hmpc_lsim_step_f64",
           "hmpc_dewh_sim_step_f64", "hmpc_dewh_control_model_f64", "hmpc_dewh_thermostat_f64",
           "hmpc_param_eval_f64", "hmpc_param_eval_v2_f64", "hmpc_param_eval_bytes_per_agent",
           "hmpc_aggregate_power_f64", "hmpc_aggregate_window_doubles", "hmpc_aggregate_publish_f64",
           "hmpc_aggregate_gather_f64", "hmpc_coupling_price_cost_f64", "hmpc_coupling_sums_f64",
           "hmpc_coupling_dual_step_f64", "hmpc_coupling_keep_best_f64", "hmpc_coupling_response_cost_f64",
           "hmpc_coupling_merge_f64", "hmpc_coupling_accept_f64", "hmpc_coupling_restore_f64", "hmpc_step_plan_create", "hmpc_step_plan_destroy", "hmpc_mpc_step_host_f64", "hmpc_mpc_step_host_bytes",
           "hmpc_step_plan_last_solver",
           "hmpc_fp64_peak_probe")

SOLVE_STATUS = {0: "optimal", 1: "infeasible", 2: "node_limit", 3: "iter_limit", 4: "numeric", 5: "unsupported"}
STAT_NAMES = ("nodes", "pivots", "cuts", "rows_added", "max_rows", "lp_solves", "purges", "reserved")
DP_STAT_NAMES = ("nodes", "unused1", "unused2", "cells", "max_open", "incumbent_updates", "unused6", "kilo_fma")


class HmpcError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("B", "Nt", "nx", "nu", "ndelta", "nz", "nmu", "nomega", "ny", "nc")]

    @property
    def nv(self):
        return self.nu + self.ndelta + self.nz + self.nmu


class MilpOpts(C.Structure):
    _fields_ = [("mip_rel_gap", C.c_double), ("int_tol", C.c_double), ("feas_tol", C.c_double),
                ("big_bound", C.c_double), ("max_nodes", C.c_int32), ("max_pivots", C.c_int32),
                ("max_cuts", C.c_int32), ("max_rows", C.c_int32), ("cut_rounds_root", C.c_int32),
                ("cut_rounds_node", C.c_int32), ("cuts_per_round", C.c_int32), ("force_general", C.c_int32)]


class StageDpOpts(C.Structure):
    _fields_ = [("mip_rel_gap", C.c_double), ("feas_tol", C.c_double), ("cells", C.c_int32), ("max_nodes", C.c_int32),
                ("table_fp64", C.c_int32), ("bound", C.c_int32), ("fuse_search", C.c_int32), ("reserved", C.c_int32)]


class MiqpOpts(C.Structure):
    _fields_ = [("mip_rel_gap", C.c_double), ("int_tol", C.c_double), ("eps", C.c_double), ("rho", C.c_double),
                ("max_nodes", C.c_int32), ("max_iter", C.c_int32)]


class StageTerms(C.Structure):
    _fields_ = [("T", C.c_int32), ("reserved", C.c_int32), ("h", C.c_void_p), ("h_stride_b", C.c_int64),
                ("ga", C.c_void_p), ("ga_stride_b", C.c_int64), ("r", C.c_void_p), ("wq", C.c_void_p),
                ("wq_stride_b", C.c_int64), ("w1", C.c_void_p), ("w1_stride_b", C.c_int64), ("qmu", C.c_void_p),
                ("qmu_stride_b", C.c_int64)]


if not os.path.exists(LIB_PATH):
    raise ImportError("libhmpc.so is not built (%s).  Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                      "`make -C pyhybridcontrol_b200/csrc`.  There is no CPU fallback." % LIB_PATH)

_lib = C.CDLL(LIB_PATH)
_P = C.c_void_p
_MatArr = _P * len(MAT_NAMES)
_StrideArr = C.c_int64 * len(MAT_NAMES)
_EvoArr = _P * len(EVO_NAMES)

_lib.hmpc_version.restype = C.c_int
_lib.hmpc_last_cuda_error.restype = C.c_char_p
_lib.hmpc_device_info.argtypes = [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_size_t)]
_lib.hmpc_condense_f64.argtypes = [C.POINTER(Dims), _MatArr, _StrideArr, _EvoArr, _P]
_lib.hmpc_condense_bytes_per_agent.argtypes = [C.POINTER(Dims)]
_lib.hmpc_condense_bytes_per_agent.restype = C.c_int64
_lib.hmpc_constraint_rhs_f64.argtypes = [C.POINTER(Dims), C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, _P]
_lib.hmpc_predict_f64.argtypes = [C.c_int32] * 5 + [_P] * 9
_lib.hmpc_linear_cost_f64.argtypes = [C.c_int32] * 4 + [_P, C.c_int64] + [_P] * 9
_lib.hmpc_milp_default_opts.argtypes = [C.POINTER(MilpOpts)]
_lib.hmpc_milp_default_opts.restype = None
_lib.hmpc_milp_workspace_bytes.argtypes = [C.c_int32] * 3 + [C.POINTER(MilpOpts), C.POINTER(C.c_size_t)]
_lib.hmpc_milp_solve_f64.argtypes = [C.c_int32] * 3 + [_P, C.c_int64, _P, C.c_int64, _P, _P, _P, C.c_int64, _P,
                                                       C.POINTER(MilpOpts), _P, C.c_size_t, _P, _P, _P, _P, _P]
_lib.hmpc_miqp_default_opts.argtypes = [C.POINTER(MiqpOpts)]
_lib.hmpc_miqp_default_opts.restype = None
_lib.hmpc_miqp_workspace_bytes.argtypes = [C.c_int32] * 3 + [C.POINTER(C.c_size_t)]
_lib.hmpc_miqp_solve_f64.argtypes = [C.c_int32] * 3 + [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, _P, _P, _P,
                                                       C.POINTER(MiqpOpts), _P, C.c_size_t, _P, _P, _P, _P, _P]
_lib.hmpc_stage_dp_default_opts.argtypes = [C.POINTER(StageDpOpts)]
_lib.hmpc_stage_dp_default_opts.restype = None
_lib.hmpc_stage_dp_supported.argtypes = [C.POINTER(Dims)]
_lib.hmpc_stage_dp_workspace_bytes.argtypes = [C.POINTER(Dims), C.POINTER(StageDpOpts), C.POINTER(C.c_size_t)]
_lib.hmpc_stage_dp_max_cells.argtypes = [C.POINTER(Dims), C.POINTER(StageDpOpts), C.POINTER(C.c_int32)]
_lib.hmpc_stage_dp_solve_f64.argtypes = [C.POINTER(Dims), _MatArr, _StrideArr, _P, _P, C.c_int64, _P, _P, _P,
                                         C.POINTER(StageTerms), C.POINTER(StageDpOpts), _P, C.c_size_t, _P, _P, _P,
                                         _P, _P]
_lib.hmpc_lsim_step_f64.argtypes = [C.POINTER(Dims), _MatArr, _StrideArr] + [_P] * 5 + [C.c_double] + [_P] * 4
_lib.hmpc_dewh_sim_step_f64.argtypes = [C.c_int32] + [_P] * 8
_lib.hmpc_dewh_control_model_f64.argtypes = [C.c_int32, _P, _P, _P]
_lib.hmpc_dewh_thermostat_f64.argtypes = [C.c_int32, _P, _P, C.c_int64, _P, _P, _P, _P]
_lib.hmpc_param_eval_f64.argtypes = [C.c_int32] * 4 + [_P, C.c_int32, C.POINTER(C.c_int32), _P, _P, _P]
_lib.hmpc_param_eval_v2_f64.argtypes = _lib.hmpc_param_eval_f64.argtypes
_lib.hmpc_param_eval_bytes_per_agent.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
_lib.hmpc_param_eval_bytes_per_agent.restype = C.c_int64
_lib.hmpc_aggregate_power_f64.argtypes = [C.c_int32, C.c_int32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P]
_lib.hmpc_coupling_price_cost_f64.argtypes = [C.c_int32] * 4 + [_P, _P, _P, C.c_int64, _P]
_lib.hmpc_coupling_sums_f64.argtypes = [C.c_int32, C.c_int32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]
_lib.hmpc_coupling_dual_step_f64.argtypes = [C.c_int32, _P, _P, _P, _P, _P, C.c_double, _P, _P, _P, _P]
_lib.hmpc_coupling_keep_best_f64.argtypes = [C.c_int32, C.c_int32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]
_lib.hmpc_coupling_response_cost_f64.argtypes = [C.c_int32] * 4 + [_P] * 6 + [C.c_int64, _P]
_lib.hmpc_coupling_merge_f64.argtypes = [C.c_int32] * 5 + [_P] * 4 + [C.c_int64] + [_P] * 5
_lib.hmpc_coupling_accept_f64.argtypes = [C.c_int32] + [_P] * 8
_lib.hmpc_coupling_restore_f64.argtypes = [C.c_int32] * 3 + [_P] * 4 + [C.c_int64, _P, _P]
_lib.hmpc_step_plan_create.argtypes = [C.POINTER(Dims), C.POINTER(MilpOpts), C.POINTER(_P)]
_lib.hmpc_step_plan_destroy.argtypes = [_P]
_lib.hmpc_mpc_step_host_f64.argtypes = [_P, C.c_int32, _MatArr, _StrideArr, _P, _P, _P, C.c_int64, _P, _P, _P,
                                        _P, _P, _P, _P, _P]
_lib.hmpc_mpc_step_host_bytes.argtypes = [_P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
_lib.hmpc_step_plan_last_solver.argtypes = [_P]
_lib.hmpc_fp64_peak_probe.argtypes = [C.POINTER(C.c_double), _P]

# number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0


def lib():
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = _lib.hmpc_last_cuda_error().decode() if rc == -2 else ""
        raise HmpcError("%s failed with status %d %s" % (what, rc, msg))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device buffers must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def default_opts(**kw):
    o = MilpOpts()
    _lib.hmpc_milp_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def make_dims(B, Nt, nx=0, nu=0, ndelta=0, nz=0, nmu=0, nomega=0, ny=0, nc=0):
    return Dims(B, Nt, nx, nu, ndelta, nz, nmu, nomega, ny, nc)


def _mat_shape(d, name):
    rows = d.nx if name in MAT_NAMES[0:6] else (d.ny if name in MAT_NAMES[6:12] else d.nc)
    cols = {"A": d.nx, "C": d.nx, "E": d.nx, "B1": d.nu, "D1": d.nu, "F1": d.nu, "B2": d.ndelta, "D2": d.ndelta,
            "F2": d.ndelta, "B3": d.nz, "D3": d.nz, "F3": d.nz, "B4": d.nomega, "D4": d.nomega, "F4": d.nomega,
            "G": d.ny, "Psi": d.nmu}.get(name, 1)
    return rows, cols


def _pack_mats(d, mats):
    arr, strides, keep = _MatArr(), _StrideArr(), []
    for i, name in enumerate(MAT_NAMES):
        t = mats.get(name)
        r, c = _mat_shape(d, name)
        if t is None or r * c == 0:
            arr[i], strides[i] = None, 0
            continue
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if tuple(t.shape[1:]) != (r, c) or t.shape[0] not in (1, d.B) or t.dtype != torch.float64:
            raise ValueError("matrix %s: expected [%d or 1, %d, %d] float64, got %s %s" % (name, d.B, r, c,
                                                                                      tuple(t.shape), t.dtype))
        t = t.contiguous()
        keep.append(t)
        arr[i] = t.data_ptr()
        strides[i] = 0 if t.shape[0] == 1 and d.B != 1 else r * c
    return arr, strides, keep


def evo_shapes(d):
    nvt, nwt = d.nv * d.Nt, d.nomega * d.Nt
    out = {}
    for grp, rows in (("Phi_x Gamma_v Gamma_omega Gamma_5", d.nx), ("L_x L_v L_omega L_5", d.ny),
                      ("H_x H_v H_omega H_5", d.nc)):
        for name, cols in zip(grp.split(), (d.nx, nvt, nwt, 1)):
            out[name] = (d.B, rows * d.Nt, cols)
    return out


def condense(d, mats, want=EVO_NAMES, out=None):
    """K1.  mats: name -> CUDA float64 tensor [B|1, r, c].  Returns name -> tensor [B, rows*Nt, cols]."""
    global launch_count
    arr, strides, keep = _pack_mats(d, mats)
    dev = next(iter(keep)).device if keep else torch.device("cuda")
    shapes = evo_shapes(d)
    res = {} if out is None else out
    evo = _EvoArr()
    for i, name in enumerate(EVO_NAMES):
        if name in want:
            if name not in res:
                res[name] = torch.empty(shapes[name], dtype=torch.float64, device=dev)
            evo[i] = res[name].data_ptr() if res[name].numel() else None
        else:
            evo[i] = None
    _check(_lib.hmpc_condense_f64(C.byref(d), arr, strides, evo, _stream()), "hmpc_condense_f64")
    launch_count += 1
    return res


def condense_bytes_per_agent(d):
    return int(_lib.hmpc_condense_bytes_per_agent(C.byref(d)))


def constraint_rhs(d, evo, x0, w, scenarios=None, rows=None, out=None):
    """K2.  w: [B, nomega*Nt]; scenarios: [B, nomega*Nt, S] (robust row-min form)."""
    global launch_count
    rows = d.nc * d.Nt if rows is None else rows
    S = 0 if scenarios is None else scenarios.shape[2]
    ww = w if scenarios is None else scenarios
    if out is None:
        out = torch.empty((d.B, rows), dtype=torch.float64, device=evo["H_5"].device)
    _check(_lib.hmpc_constraint_rhs_f64(C.byref(d), rows, _ptr(evo["H_x"]) if d.nx else None,
                                        _ptr(evo["H_omega"]) if d.nomega else None, _ptr(evo["H_5"]),
                                        _ptr(x0) if d.nx else None, _ptr(ww) if d.nomega else None, S, _ptr(out),
                                        _stream()), "hmpc_constraint_rhs_f64")
    launch_count += 1
    return out


def predict(M_x, M_v, M_w, M_5, x0, v, w):
    """Affine prediction rows: out[b] = M_x x0 + M_v v + M_w w + M_5."""
    global launch_count
    B, R = M_5.shape[0], M_5.shape[1]
    nx = M_x.shape[2] if M_x is not None and M_x.numel() else 0
    nvt = M_v.shape[2] if M_v is not None else 0
    nwt = M_w.shape[2] if M_w is not None and M_w.numel() else 0
    out = torch.empty((B, R), dtype=torch.float64, device=M_5.device)
    _check(_lib.hmpc_predict_f64(B, R, nx, nvt, nwt, _ptr(M_x) if nx else None, _ptr(M_v), _ptr(M_w) if nwt else None,
                                 _ptr(M_5), _ptr(x0) if nx else None, _ptr(v), _ptr(w) if nwt else None, _ptr(out),
                                 _stream()), "hmpc_predict_f64")
    launch_count += 1
    return out


def linear_cost(B, nvt, w_v=None, w_x=None, Gamma_v=None, xc=None, w_y=None, L_v=None, yc=None):
    global launch_count
    ref = next(t for t in (w_v, w_x, w_y) if t is not None)
    c = torch.empty((B, nvt), dtype=torch.float64, device=ref.device)
    c0 = torch.zeros((B,), dtype=torch.float64, device=ref.device)
    stride = 0
    if w_v is not None:
        w_v = w_v.reshape(-1, nvt)
        stride = nvt if w_v.shape[0] == B and B > 1 else (0 if w_v.shape[0] == 1 and B > 1 else nvt)
    nxt = w_x.shape[1] if w_x is not None else 0
    nyt = w_y.shape[1] if w_y is not None else 0
    _check(_lib.hmpc_linear_cost_f64(B, nvt, nxt, nyt, _ptr(w_v), stride, _ptr(w_x), _ptr(Gamma_v), _ptr(xc),
                                     _ptr(w_y), _ptr(L_v), _ptr(yc), _ptr(c), _ptr(c0), _stream()),
           "hmpc_linear_cost_f64")
    launch_count += 1
    return c, c0


def milp_solve(c, H, rhs, lb, ub, is_bin, opts=None):
    """K3/K4.  c [B|1,n], H [B|1,m,n], rhs [B,m], lb/ub [B|1,n] (or [n]), is_bin uint8 [n]."""
    global launch_count
    B, m = rhs.shape
    n = c.shape[-1]
    c2 = c.reshape(-1, n)
    H3 = H.reshape(-1, m, n) if m else H
    lb2, ub2 = lb.reshape(-1, n), ub.reshape(-1, n)
    dev = rhs.device
    v = torch.empty((B, n), dtype=torch.float64, device=dev)
    obj = torch.empty((B,), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    stats = torch.empty((B, 8), dtype=torch.int32, device=dev)
    o = opts if opts is not None else default_opts()
    _check(_lib.hmpc_milp_solve_f64(B, n, m, _ptr(c2), n if c2.shape[0] == B and B > 1 or B == 1 else 0,
                                    _ptr(H3) if m else None, (m * n) if (m and H3.shape[0] == B) else 0,
                                    _ptr(rhs) if m else None, _ptr(lb2), _ptr(ub2),
                                    n if lb2.shape[0] == B and B > 1 else 0, _ptr(is_bin), C.byref(o), None, 0,
                                    _ptr(v), _ptr(obj), _ptr(status), _ptr(stats), _stream()), "hmpc_milp_solve_f64")
    launch_count += 1
    return v, obj, status, stats


def miqp_default_opts(**kw):
    o = MiqpOpts()
    _lib.hmpc_miqp_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


_qp_workspaces = {}


def miqp_solve(c, H, rhs, lb, ub, is_bin, P=None, opts=None):
    """K3q/K4q.  c [B|1,n], H [B|1,m,n], rhs [B,m], lb/ub [n], is_bin uint8 [n], P [B|1,n,n] or None
    -> (v [B,n], obj [B], status [B], stats [B,8])."""
    global launch_count
    dev = rhs.device
    B, m = rhs.shape
    n = c.shape[-1]
    c2 = c.reshape(-1, n).contiguous()
    H3 = H.reshape(-1, m, n).contiguous()
    P3 = None if P is None else P.reshape(-1, n, n).contiguous()
    o = opts if opts is not None else miqp_default_opts()
    need = C.c_size_t()
    _check(_lib.hmpc_miqp_workspace_bytes(B, n, m, C.byref(need)), "hmpc_miqp_workspace_bytes")
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _qp_workspaces.get(key)
    if ws is None or ws.numel() < need.value:
        ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
        _qp_workspaces[key] = ws
    v = torch.empty((B, n), dtype=torch.float64, device=dev)
    obj = torch.empty((B,), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    stats = torch.empty((B, 8), dtype=torch.int32, device=dev)
    _check(_lib.hmpc_miqp_solve_f64(B, n, m, None if P3 is None else _ptr(P3), 0 if (P3 is None or P3.shape[0] == 1 and B > 1) else n * n,
                                    _ptr(c2), 0 if (c2.shape[0] == 1 and B > 1) else n,
                                    _ptr(H3), 0 if (H3.shape[0] == 1 and B > 1) else m * n, _ptr(rhs.contiguous()),
                                    _ptr(lb.reshape(-1)), _ptr(ub.reshape(-1)), _ptr(is_bin), C.byref(o),
                                    C.c_void_p(ws.data_ptr()), need.value, _ptr(v), _ptr(obj), _ptr(status), _ptr(stats),
                                    _stream()), "hmpc_miqp_solve_f64")
    launch_count += 1
    return v, obj, status, stats


def stage_dp_default_opts(**kw):
    o = StageDpOpts()
    _lib.hmpc_stage_dp_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


DP_BOUND_CONSTANT, DP_BOUND_LINEAR = 0, 1


def stage_dp_max_cells(d, opts=None):
    """Largest cell count whose two stage buffers fit shared memory for these dimensions / this cell format."""
    o = opts if opts is not None else stage_dp_default_opts()
    n = C.c_int32()
    _check(_lib.hmpc_stage_dp_max_cells(C.byref(d), C.byref(o), C.byref(n)), "hmpc_stage_dp_max_cells")
    return int(n.value)


def stage_dp_supported(d):
    """True when the DIMENSIONS fit the scalar-state class of hmpc_stage_dp_solve_f64 (the matrix pattern --
    Psi = -diag(d), A > 0 -- is checked per agent on the device and reported as status 5)."""
    return bool(_lib.hmpc_stage_dp_supported(C.byref(d)))


_dp_workspaces = {}


def _dp_workspace(d, o, dev):
    need = C.c_size_t()
    _check(_lib.hmpc_stage_dp_workspace_bytes(C.byref(d), C.byref(o), C.byref(need)), "hmpc_stage_dp_workspace_bytes")
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _dp_workspaces.get(key)
    if ws is None or ws.numel() < need.value:
        ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
        _dp_workspaces[key] = ws
    return ws, need.value


def _stage_terms(d, terms, dev):
    """dict(h [B|1,T], ga [B|1,T,nb], r [B,Nt,T], wq/w1 [B|1,Nt,T], qmu [B|1,Nt,nc]) of device tensors -> struct."""
    st = StageTerms()
    keep = []
    if not terms:
        return None, keep
    nb = d.nu + d.ndelta
    T = 0 if terms.get("h") is None else terms["h"].shape[-1]
    st.T = T

    def put(name, inner):
        t = terms.get(name)
        if t is None:
            return None, 0
        t = t.to(device=dev, dtype=torch.float64).reshape((-1,) + inner).contiguous()
        keep.append(t)
        n = int(np.prod(inner))
        return t.data_ptr(), (n if t.shape[0] == d.B and d.B > 1 or d.B == 1 else 0)
    if T:
        st.h, st.h_stride_b = put("h", (T,))
        st.ga, st.ga_stride_b = put("ga", (T, nb))
        r = terms["r"].to(device=dev, dtype=torch.float64).reshape(d.B, d.Nt, T).contiguous()
        keep.append(r)
        st.r = r.data_ptr()
        st.wq, st.wq_stride_b = put("wq", (d.Nt, T))
        st.w1, st.w1_stride_b = put("w1", (d.Nt, T))
    st.qmu, st.qmu_stride_b = put("qmu", (d.Nt, d.nc))
    return st, keep


def stage_dp_launches(B, fuse_search=-1):
    """kernels one hmpc_stage_dp_solve_f64 call launches: the table kernel alone when the search is its tail (up to 296
    agents by default), else table + one-warp search + team search"""
    fused = (int(B) <= 296) if int(fuse_search) < 0 else bool(fuse_search)
    return 1 if fused else 3


def stage_dp_solve(d, mats, rhs, cost_v, lb, ub, is_bin, opts=None, terms=None):
    """K3s/K4s.  mats as for condense(); rhs [B, nc*Nt]; cost_v [B|1, nv*Nt]; lb/ub [nv*Nt]; is_bin uint8 [nv*Nt];
    terms: optional convex state / slack terms (see hmpc_stage_terms) -> an MIQP."""
    global launch_count
    arr, strides, keep = _pack_mats(d, mats)
    nvt = (d.nu + d.ndelta + d.nmu) * d.Nt
    dev = cost_v.device
    c2 = cost_v.reshape(-1, nvt)
    o = opts if opts is not None else stage_dp_default_opts()
    ws, nbytes = _dp_workspace(d, o, dev)
    st, keep_terms = _stage_terms(d, terms, dev)
    v = torch.empty((d.B, nvt), dtype=torch.float64, device=dev)
    obj = torch.empty((d.B,), dtype=torch.float64, device=dev)
    status = torch.empty((d.B,), dtype=torch.int32, device=dev)
    stats = torch.empty((d.B, 8), dtype=torch.int32, device=dev)
    _check(_lib.hmpc_stage_dp_solve_f64(C.byref(d), arr, strides, _ptr(rhs) if d.nc else None, _ptr(c2),
                                        nvt if (c2.shape[0] == d.B and d.B > 1) or d.B == 1 else 0,
                                        _ptr(lb.reshape(-1)), _ptr(ub.reshape(-1)), _ptr(is_bin),
                                        C.byref(st) if st is not None else None, C.byref(o),
                                        C.c_void_p(ws.data_ptr()), nbytes, _ptr(v), _ptr(obj), _ptr(status),
                                        _ptr(stats), _stream()), "hmpc_stage_dp_solve_f64")
    launch_count += stage_dp_launches(d.B, o.fuse_search)
    return v, obj, status, stats


def lsim_step(d, mats, x, u, delta, z, w, cons_tol=1e-6):
    """K5 generic MLD step -> (x1 [B,nx], y [B,ny], cons uint8 [B,nc])."""
    global launch_count
    arr, strides, keep = _pack_mats(d, mats)
    dev = next(t for t in (x, u, delta, z, w) if t is not None).device
    x1 = torch.empty((d.B, d.nx), dtype=torch.float64, device=dev)
    y = torch.empty((d.B, d.ny), dtype=torch.float64, device=dev)
    cons = torch.empty((d.B, d.nc), dtype=torch.uint8, device=dev)
    _check(_lib.hmpc_lsim_step_f64(C.byref(d), arr, strides, _ptr(x) if d.nx else None, _ptr(u) if d.nu else None,
                                   _ptr(delta) if d.ndelta else None, _ptr(z) if d.nz else None,
                                   _ptr(w) if d.nomega else None, cons_tol, _ptr(x1) if d.nx else None,
                                   _ptr(y) if d.ny else None, _ptr(cons) if d.nc else None, _stream()),
           "hmpc_lsim_step_f64")
    launch_count += (1 if d.nx + d.ny else 0) + (1 if d.nc else 0)
    return x1, y, cons


DEWH_PARAM_ORDER = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom", "ts")


def pack_dewh_params(param_list):
    """list of param dicts -> float64 numpy [B,12] in the order hmpc_dewh_* expects."""
    out = np.zeros((len(param_list), 12))
    for b, p in enumerate(param_list):
        out[b, :11] = [p[k] for k in DEWH_PARAM_ORDER]
    return out


def dewh_sim_step(params, T, u, D_h, want_model=False):
    global launch_count
    B = T.shape[0]
    T1 = torch.empty_like(T)
    model = torch.empty((B, 4), dtype=torch.float64, device=T.device) if want_model else None
    cons = torch.empty((B, 2), dtype=torch.uint8, device=T.device)
    _check(_lib.hmpc_dewh_sim_step_f64(B, _ptr(params), _ptr(T), _ptr(u), _ptr(D_h), _ptr(T1), _ptr(model),
                                       _ptr(cons), _stream()), "hmpc_dewh_sim_step_f64")
    launch_count += 1
    return T1, model, cons


def dewh_thermostat(params, band, T, u_prev):
    """Thermostat rule for a batch: band [B,2] or [1,2] = (T_h_max_sub_T_h_on, T_h_max_sub_T_h_off) -> u [B]."""
    global launch_count
    B = T.shape[0]
    band = band.reshape(-1, 2)
    if band.shape[0] not in (1, B):
        raise ValueError("band must be [B,2] or [1,2]")
    u = torch.empty_like(T)
    band = band.contiguous()
    _check(_lib.hmpc_dewh_thermostat_f64(B, _ptr(params), _ptr(band), 2 if band.shape[0] == B else 0, _ptr(T),
                                         _ptr(u_prev), _ptr(u), _stream()), "hmpc_dewh_thermostat_f64")
    launch_count += 1
    return u


def dewh_control_model(params):
    global launch_count
    B = params.shape[0]
    model = torch.empty((B, 4), dtype=torch.float64, device=params.device)
    _check(_lib.hmpc_dewh_control_model_f64(B, _ptr(params), _ptr(model), _stream()), "hmpc_dewh_control_model_f64")
    launch_count += 1
    return model


def param_eval(program, n_regs, mat_sizes, params, out=None, version=1):
    """Symbolic / callable model front-end (hmpc.h: hmpc_param_eval_f64).  program: CUDA int32 tensor [n_ins, 4];
    params: CUDA float64 [B, P]; returns the flat output buffer, matrix m at [B*off_m, B*(off_m+size_m)) as
    [B, size_m].  version=2: the experimental kernel (parameters preloaded as registers 0..P-1, two agents per
    thread; hmpc_param_eval_v2_f64) -- the program must be in that form (ExprProgram.instructions_v2)."""
    global launch_count
    if program.dtype != torch.int32 or program.dim() != 2 or program.shape[1] != 4:
        raise ValueError("program must be an int32 tensor [n_ins, 4]")
    if params.dtype != torch.float64 or params.dim() != 2:
        raise ValueError("params must be a float64 tensor [B, P]")
    B, P = params.shape
    sizes = (C.c_int32 * len(mat_sizes))(*[int(s) for s in mat_sizes])
    n_out = int(sum(mat_sizes))
    if out is None:
        out = torch.empty((B * n_out,), dtype=torch.float64, device=params.device)
    elif out.numel() != B * n_out or out.dtype != torch.float64:
        raise ValueError("out must hold B * sum(mat_sizes) doubles")
    fn = _lib.hmpc_param_eval_f64 if int(version) == 1 else _lib.hmpc_param_eval_v2_f64
    _check(fn(B, P, int(n_regs), program.shape[0], _ptr(program), len(mat_sizes), sizes,
              _ptr(params) if P else None, _ptr(out), _stream()), "hmpc_param_eval_f64 (version %d)" % int(version))
    launch_count += 1
    return out


def param_eval_bytes_per_agent(n_params, mat_sizes):
    sizes = (C.c_int32 * len(mat_sizes))(*[int(s) for s in mat_sizes])
    return int(_lib.hmpc_param_eval_bytes_per_agent(int(n_params), len(mat_sizes), sizes))


COUPLING_STATE = ("lower_bound", "upper_bound", "dual", "primal", "improved", "iterations", "gnorm2", "skipped")


def coupling_state(device):
    """fresh state vector of the price coordination (hmpc.h: hmpc_coupling_dual_step_f64)"""
    return torch.tensor([-float("inf"), float("inf"), 0, 0, 0, 0, 0, 0], dtype=torch.float64, device=device)


def coupling_price_cost(lam, P_nom, cost_v, nv, col=0):
    """cost_v [B, Nt*nv] <- lambda_k * P_nom_b in column `col` of every step (in place)."""
    global launch_count
    B, Nt = cost_v.shape[0], lam.shape[0]
    _check(_lib.hmpc_coupling_price_cost_f64(B, Nt, nv, col, _ptr(lam), _ptr(P_nom), _ptr(cost_v), cost_v.stride(0),
                                             _stream()), "hmpc_coupling_price_cost_f64")
    launch_count += 1


def coupling_sums(u, P_nom, obj, status, sums):
    """sums [Nt+2] <- {sum_b P_nom[b] u[b,k], sum_b obj[b], #(status != 0)}; u [B, Nt] with any strides."""
    global launch_count
    B, Nt = u.shape
    _check(_lib.hmpc_coupling_sums_f64(B, Nt, C.c_void_p(u.data_ptr()), u.stride(0), u.stride(1), _ptr(P_nom), _ptr(obj),
                                       _ptr(status), _ptr(sums), _stream()), "hmpc_coupling_sums_f64")
    launch_count += 1


def coupling_dual_step(sums, p_other, price, a_lo, a_hi, theta, lam, lam_next, state):
    global launch_count
    _check(_lib.hmpc_coupling_dual_step_f64(lam.shape[0], _ptr(sums), _ptr(p_other), _ptr(price), _ptr(a_lo), _ptr(a_hi),
                                            float(theta), _ptr(lam), _ptr(lam_next), _ptr(state), _stream()),
           "hmpc_coupling_dual_step_f64")
    launch_count += 1


def coupling_keep_best(u, lam, state, u_best, lam_best):
    global launch_count
    B, Nt = u.shape
    _check(_lib.hmpc_coupling_keep_best_f64(B, Nt, C.c_void_p(u.data_ptr()), u.stride(0), u.stride(1), _ptr(lam),
                                            _ptr(state), _ptr(u_best), _ptr(lam_best), _stream()),
           "hmpc_coupling_keep_best_f64")
    launch_count += 1


def coupling_response_cost(agg, v_cur, P_nom, p_other, price, cost_v, nv, col=0):
    """cost_v[b, k*nv+col] <- agent b's marginal import price at step k with the others fixed (in place)."""
    global launch_count
    B, Nt = v_cur.shape[0], price.shape[0]
    assert v_cur.stride(0) == cost_v.stride(0)
    _check(_lib.hmpc_coupling_response_cost_f64(B, Nt, nv, col, _ptr(agg), _ptr(v_cur), _ptr(P_nom), _ptr(p_other),
                                                _ptr(price), _ptr(cost_v), cost_v.stride(0), _stream()),
           "hmpc_coupling_response_cost_f64")
    launch_count += 1


def coupling_merge(lo, hi, Nt, nv, v_new, obj_new, status_new, cost_v, v_cur, pen_cur, v_bak, pen_bak, col=0):
    global launch_count
    assert v_cur.stride(0) == cost_v.stride(0)
    _check(_lib.hmpc_coupling_merge_f64(lo, hi, Nt, nv, col, _ptr(v_new), _ptr(obj_new), _ptr(status_new), _ptr(cost_v),
                                        cost_v.stride(0), _ptr(v_cur), _ptr(pen_cur), _ptr(v_bak), _ptr(pen_bak),
                                        _stream()), "hmpc_coupling_merge_f64")
    launch_count += 1


def coupling_accept(sums_cand, p_other, price, a_lo, a_hi, sums_cur, br_state):
    global launch_count
    _check(_lib.hmpc_coupling_accept_f64(price.shape[0], _ptr(sums_cand), _ptr(p_other), _ptr(price), _ptr(a_lo),
                                         _ptr(a_hi), _ptr(sums_cur), _ptr(br_state), _stream()),
           "hmpc_coupling_accept_f64")
    launch_count += 1


def coupling_restore(lo, hi, br_state, v_bak, pen_bak, v_cur, pen_cur):
    global launch_count
    _check(_lib.hmpc_coupling_restore_f64(lo, hi, v_cur.shape[1], _ptr(br_state), _ptr(v_bak), _ptr(pen_bak),
                                          _ptr(v_cur), v_cur.stride(0), _ptr(pen_cur), _stream()),
           "hmpc_coupling_restore_f64")
    launch_count += 1


def aggregate_power(u, P_nom=None):
    """K6 local part: u [B, Nt] (any strides) -> P_agg [Nt] = sum_b P_nom[b] u[b,k]."""
    global launch_count
    B, Nt = u.shape
    chunks = max(1, (B + 15) // 16)
    partial = torch.empty((chunks, Nt), dtype=torch.float64, device=u.device)
    out = torch.empty((Nt,), dtype=torch.float64, device=u.device)
    _check(_lib.hmpc_aggregate_power_f64(B, Nt, C.c_void_p(u.data_ptr()), u.stride(0), u.stride(1), _ptr(P_nom),
                                         _ptr(partial), _ptr(out), _stream()), "hmpc_aggregate_power_f64")
    launch_count += 2
    return out


_lib.hmpc_aggregate_window_doubles.argtypes = [C.c_int32, C.c_int32]
_lib.hmpc_aggregate_window_doubles.restype = C.c_int64


def aggregate_window_doubles(Nt, world):
    return int(_lib.hmpc_aggregate_window_doubles(int(Nt), int(world)))


def aggregate_publish(u, P_nom, world, rank, windows_dev, out_prev=None, lag=1):
    """K6 local reduction + publication into every rank's exchange window (windows_dev: int64 CUDA tensor of `world`
    device pointers)."""
    global launch_count
    B, Nt = u.shape
    chunks = max(1, (B + 15) // 16)
    partial = torch.empty((chunks, Nt), dtype=torch.float64, device=u.device)
    _check(_lib.hmpc_aggregate_publish_f64(B, Nt, C.c_void_p(u.data_ptr()), C.c_int64(u.stride(0)), C.c_int32(u.stride(1)),
                                           _ptr(P_nom), _ptr(partial), C.c_int32(world), C.c_int32(rank),
                                           C.c_void_p(windows_dev.data_ptr()), _ptr(out_prev), C.c_int32(lag), _stream()),
           "hmpc_aggregate_publish_f64")
    launch_count += 2
    return partial


def aggregate_gather(Nt, world, rank, window, out=None, spin_limit=0, lag=0):
    global launch_count
    out = out if out is not None else torch.empty((Nt,), dtype=torch.float64, device=window.device)
    _check(_lib.hmpc_aggregate_gather_f64(C.c_int32(Nt), C.c_int32(world), C.c_int32(rank), C.c_void_p(window.data_ptr()),
                                          _ptr(out), C.c_int64(spin_limit), C.c_int32(lag), _stream()),
           "hmpc_aggregate_gather_f64")
    launch_count += 1
    return out


def fp64_peak_tflops():
    v = C.c_double(0.0)
    _check(_lib.hmpc_fp64_peak_probe(C.byref(v), _stream()), "hmpc_fp64_peak_probe")
    return v.value


def device_info():
    sm, ma, mi, sh = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    rc = _lib.hmpc_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sh))
    return dict(rc=rc, sm_count=sm.value, cc=(ma.value, mi.value), smem_optin=sh.value)


class StepPlan(object):
    """Host-buffer front door (hmpc_mpc_step_host_f64): numpy in, numpy out, copies inside the call."""

    def __init__(self, d, opts=None):
        self.d = d
        self.opts = opts if opts is not None else default_opts()
        self._h = _P()
        _check(_lib.hmpc_step_plan_create(C.byref(d), C.byref(self.opts), C.byref(self._h)), "hmpc_step_plan_create")
        nvt = d.nv * d.Nt
        self.v = np.empty((d.B, nvt))
        self.obj = np.empty((d.B,))
        self.status = np.empty((d.B,), dtype=np.int32)
        self.stats = np.empty((d.B, 8), dtype=np.int32)
        self.timing = (C.c_float * 4)()
        self._out_ptrs = (self.v.ctypes.data, self.obj.ctypes.data, self.status.ctypes.data, self.stats.ctypes.data)
        self._mats_cache = {}

    def bytes_per_step(self, recondense):
        a, b = C.c_int64(), C.c_int64()
        _check(_lib.hmpc_mpc_step_host_bytes(self._h, int(recondense), C.byref(a), C.byref(b)), "bytes")
        return a.value, b.value

    def _pack_host_mats(self, mats):
        """ctypes views of the host MLD blocks; cached per dict object (the arrays are read at call time, so in-place
        updates of their contents are picked up)."""
        key = id(mats)
        hit = self._mats_cache.get(key)
        if hit is not None and hit[0] is mats:
            return hit[1], hit[2]
        d = self.d
        arr, strides, keep = _MatArr(), _StrideArr(), []
        copied = False
        for i, name in enumerate(MAT_NAMES):
            a = None if mats is None else mats.get(name)
            r, c = _mat_shape(d, name)
            if a is None or r * c == 0:
                arr[i], strides[i] = None, 0
                continue
            copied |= not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous)
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, r, c)
            keep.append(a)
            arr[i] = a.ctypes.data
            strides[i] = 0 if a.shape[0] == 1 and d.B != 1 else r * c
        self._mats_cache = {} if copied else {key: (mats, arr, strides, keep)}   # converted copies would go stale
        self._mats_keep = keep
        return arr, strides

    @staticmethod
    def _f64(a):
        return a if (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous) else \
            np.ascontiguousarray(a, dtype=np.float64)

    def step(self, mats, x0, w, cost_v, lb_v, ub_v, is_bin_v, recondense=True):
        global launch_count
        d = self.d
        arr, strides = self._pack_host_mats(mats)
        nvt = d.nv * d.Nt
        x0, w, lb_v, ub_v = self._f64(x0), self._f64(w), self._f64(lb_v), self._f64(ub_v)
        cost_v = self._f64(cost_v).reshape(-1, nvt)
        if not (isinstance(is_bin_v, np.ndarray) and is_bin_v.dtype == np.uint8 and is_bin_v.flags.c_contiguous):
            is_bin_v = np.ascontiguousarray(is_bin_v, dtype=np.uint8)
        cs = nvt if cost_v.shape[0] == d.B and d.B > 1 else (0 if d.B > 1 else nvt)
        _check(_lib.hmpc_mpc_step_host_f64(self._h, int(recondense), arr, strides, x0.ctypes.data, w.ctypes.data,
                                           cost_v.ctypes.data, cs, lb_v.ctypes.data, ub_v.ctypes.data,
                                           is_bin_v.ctypes.data, self._out_ptrs[0], self._out_ptrs[1],
                                           self._out_ptrs[2], self._out_ptrs[3], self.timing),
               "hmpc_mpc_step_host_f64")
        self.last_solver = ("bnc", "stage_dp")[max(0, _lib.hmpc_step_plan_last_solver(self._h))]
        launch_count += (1 if recondense else 0) + 1 + (stage_dp_launches(d.B) if self.last_solver == "stage_dp" else 1)
        return self.v, self.obj, self.status, self.stats, tuple(self.timing)

    def close(self):
        if self._h:
            _lib.hmpc_step_plan_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
