"""Numeric MLD blocks of the example's non-DEWH devices (reference: examples/.../modelling/micro_grid_models.py).

The DEWH model lives in the CUDA library (hmpc_dewh_control_model_f64 / hmpc_dewh_sim_step_f64); these are the
constant blocks of the grid, PV and residential-demand agents, which have no state."""
import numpy as np


def grid_mld_matrices(p, num_devices):
    """Grid agent (micro_grid_models.py:137-172): y = 1' omega, one binary delta = [y >= 0], one auxiliary
    z = delta y, six rows  F2 delta + F3 z + G y <= f5:
        y >= P_min (1 - delta),  y <= (P_max + eps) delta - eps,  z <= P_max delta,  z >= P_min delta,
        z <= y - P_min (1 - delta),  z >= y - P_max (1 - delta)."""
    lo, hi, eps = float(p["P_g_min"]), float(p["P_g_max"]), float(p["eps"])
    return dict(D4=np.ones((1, int(num_devices))),
                F2=np.array([[-lo, -(hi + eps), -hi, lo, -lo, hi]]).T,
                F3=np.array([[0.0, 0.0, 1.0, -1.0, 1.0, -1.0]]).T,
                f5=np.array([[-lo, -eps, 0.0, 0.0, -lo, hi]]).T,
                G=np.array([[-1.0, 1.0, 0.0, 0.0, -1.0, 1.0]]).T)


def pv_gain(p):
    """PV agent output y = -P_pv_max P_pv_units omega (micro_grid_models.py:175-203)."""
    return -float(p["P_pv_max"]) * float(p["P_pv_units"])


def resd_gain(p):
    """Residential demand output y = P_res_ave P_res_units omega (micro_grid_models.py:206-234)."""
    return float(p["P_res_ave"]) * float(p["P_res_units"])
