"""Physical parameter sets of the example micro-grid (reference: examples/.../modelling/parameters.py:10-69)."""
import numpy as np

control_ts_seconds = 900.0  # 15 min

dewh_param_struct = dict(
    C_w=4.1816e3,   # J/kg/K
    A_h=2.35,       # m^2
    U_h=0.88,       # W/m^2/K
    m_h=150.0,      # kg
    T_w=15.0,       # C
    T_inf=25.0,     # C
    P_h_Nom=3000.0, # W
    T_h_min=50.0,   # C
    T_h_max=65.0,   # C
    T_h_Nom=45.0,   # C
    T_h_max_sub_T_h_on=12,
    T_h_max_sub_T_h_off=4,
    T_h=45.0,       # C
    D_h=0.0,        # kg/s
    ts=control_ts_seconds,
)

grid_param_struct = dict(P_g_min=-2e4, P_g_max=2e4, eps=float(np.finfo(float).eps), ts=control_ts_seconds)
pv_param_struct = dict(P_pv_max=2000.0, P_pv_units=1, ts=control_ts_seconds)
res_demand_param_struct = dict(P_res_ave=1200.0, P_res_units=1, ts=control_ts_seconds)

# Eskom-style time-of-use levels in c/kWh (reference: examples/.../micro_grid_control_simulation.py:86-87)
TOU_LEVELS = dict(low_off_peak=48.40, low_stnd=76.28, low_peak=110.84,
                  high_off_peak=55.90, high_stnd=102.95, high_peak=339.77)
