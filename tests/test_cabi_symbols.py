"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/hmpc.h declares.
No compute calls are made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyhybridcontrol_b200", "csrc", "libhmpc.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hmpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmpc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("hmpc_condense_f64", "hmpc_constraint_rhs_f64", "hmpc_milp_solve_f64", "hmpc_lsim_step_f64",
                 "hmpc_dewh_sim_step_f64", "hmpc_aggregate_power_f64", "hmpc_mpc_step_host_f64"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), "libhmpc.so does not export %s" % name


def test_version_and_defaults(lib):
    lib.hmpc_version.restype = ctypes.c_int
    assert lib.hmpc_version() >= 100
    from pyhybridcontrol_b200 import cabi
    o = cabi.default_opts()
    assert o.mip_rel_gap == 0.0 and o.int_tol == 1e-6 and o.max_nodes > 0
    assert set(cabi.EXPORTS) == set(declared_symbols())


def test_argument_errors_without_gpu(lib):
    # pure argument validation returns HMPC_ERR_ARG before touching the device
    lib.hmpc_condense_f64.restype = ctypes.c_int
    assert lib.hmpc_condense_f64(None, None, None, None, None) == -1
    lib.hmpc_condense_bytes_per_agent.restype = ctypes.c_int64
    from pyhybridcontrol_b200 import cabi
    d = cabi.make_dims(1, 49, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
    assert cabi.condense_bytes_per_agent(d) == 310464      # DEWH at N_p = 48
    d = cabi.make_dims(1, 25, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
    assert cabi.condense_bytes_per_agent(d) == 81600
    d = cabi.make_dims(1, 97, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
    assert cabi.condense_bytes_per_agent(d) == 1210560
