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


# ---- symbolic device models (reference: examples/.../modelling/micro_grid_models.py:11-234) ----------------------
# The reference builds each device as an MldSystemModel whose matrices are sympy expressions of the physical
# parameters; here the same models go through the GPU front-end (one hmpc_param_eval_f64 launch per parameter
# change, or one launch for a whole fleet via get_mld_numeric_batch).
from ...models.mld_model import MldModel, MldSystemModel          # noqa: E402
from . import parameters as _par                                  # noqa: E402

_symbolic_cache = {}


def _cached(key, build):
    if key not in _symbolic_cache:
        _symbolic_cache[key] = build()
    return _symbolic_cache[key]


class PvMldSystemModel(MldSystemModel):
    """MldSystemModel + parameter schedules along the horizon (reference: models/mld_model.py:1167-1226)."""

    def get_mld_numeric_tilde(self, N_tilde, param_struct=None, param_struct_subset=None, schedule_params_tilde=None,
                              copy=None, **kwargs):
        if schedule_params_tilde is None:
            return [self.get_mld_numeric(param_struct, param_struct_subset=param_struct_subset,
                                         missing_param_check=False, invalid_param_check=False)] * N_tilde
        if len(schedule_params_tilde) != N_tilde:
            raise ValueError("Invalid length:'%d' for param_struct_tilde.schedule_param_tilde, length of "
                             "schedule_param_tilde must be equal to N_tilde:'%d'" % (len(schedule_params_tilde), N_tilde))
        return [self.get_mld_numeric(param_struct=param_struct, param_struct_subset=dict(schedule_params_tilde[k]),
                                     invalid_param_check=False) for k in range(N_tilde)]


class DewhModel(PvMldSystemModel):
    """Domestic electric water heater (reference: micro_grid_models.py:11-100).  One state (tank temperature), one
    binary input (element on), disturbance = hot-water draw, two soft temperature limits.

    ``const_heat=True``  : the draw is an energy demand at T_h_Nom, constant over the sample (control model);
    ``const_heat=False`` : the draw is a flow rate, so the loss coefficient depends on the present temperature T_h
                           (simulation model, re-evaluated every step, micro_grid_agents.py:389-408)."""

    def __init__(self, param_struct=None, const_heat=True, mld_numeric=None, mld_callable=None, mld_symbolic=None):
        param_struct = param_struct or _par.dewh_param_struct
        if mld_numeric is None and mld_callable is None and mld_symbolic is None:
            mld_symbolic = self.get_dewh_mld_symbolic(const_heat=const_heat)
        super(DewhModel, self).__init__(mld_numeric=mld_numeric, mld_symbolic=mld_symbolic,
                                        mld_callable=mld_callable, param_struct=param_struct)

    @staticmethod
    def get_dewh_mld_symbolic(const_heat=True, binary_input=True):
        return _cached(("dewh", bool(const_heat), bool(binary_input)),
                       lambda: DewhModel._build_symbolic(const_heat, binary_input))

    @staticmethod
    def _build_symbolic(const_heat, binary_input):
        import sympy as sp
        ts, C_w, A_h, U_h, m_h, D_h, T_w, T_inf, P_h_Nom, T_h_min, T_h_max, T_h_Nom, T_h = sp.symbols(
            "ts C_w A_h U_h m_h D_h T_w T_inf P_h_Nom T_h_min T_h_max T_h_Nom T_h")
        loss, cap = U_h * A_h, m_h * C_w                    # W/K to the ambient, J/K of the tank
        if const_heat:
            a_c = -loss / cap
            b4_c = C_w * (T_w - T_h_Nom) / cap
        else:
            mix = (T_h_Nom - T_w) / (T_h - T_w)             # tank water per unit of nominal-temperature water
            a_c = -(D_h * C_w * mix + loss) / cap
            b4_c = C_w * T_w * mix / cap
        # exact discretisation of the scalar system: A = e^{a_c ts}, input gain (A - 1)/a_c.  The reference writes
        # the gain as pinv(A_c)(e^{A_c ts} - I) (micro_grid_models.py:52-57), the same number for a 1x1 A_c.
        A = sp.exp(a_c * ts)
        gain = (A - 1) / a_c
        mats = dict(A=sp.Matrix([[A]]), B1=sp.Matrix([[gain * P_h_Nom / cap]]), B4=sp.Matrix([[gain * b4_c]]),
                    b5=sp.Matrix([[gain * loss * T_inf / cap]]))
        if binary_input:
            mats.update(E=np.array([[1.0], [-1.0]]), F1=np.zeros((2, 1)), Psi=-np.eye(2),
                        f5=sp.Matrix([[T_h_max], [-T_h_min]]))
        else:                                               # relaxed input 0 <= u <= 1 as two hard rows
            mats.update(E=np.array([[1.0], [-1.0], [0.0], [0.0]]), F1=np.array([[0.0], [0.0], [1.0], [-1.0]]),
                        Psi=np.array([[-1.0, 0.0], [0.0, -1.0], [0.0, 0.0], [0.0, 0.0]]),
                        f5=sp.Matrix([[T_h_max], [-T_h_min], [1.0], [0.0]]))
        return MldModel(mats, nu_l=1 if binary_input else 0, ts=0)


class GridModel(PvMldSystemModel):
    """Grid connection (reference: micro_grid_models.py:103-172): y = sum of the device powers, delta = [y >= 0],
    z = delta y (import), six big-M rows -- the rows of ``grid_mld_matrices`` with symbolic limits."""

    def __init__(self, param_struct=None, num_devices=None, mld_numeric=None, mld_callable=None, mld_symbolic=None):
        param_struct = param_struct or _par.grid_param_struct
        num_devices = 0 if num_devices is None else num_devices
        if not isinstance(num_devices, (int, np.integer)):
            raise ValueError("num_devices must be an integer")
        self._num_devices = int(num_devices)
        if mld_numeric is None and mld_callable is None and mld_symbolic is None:
            mld_symbolic = self.get_grid_mld_symbolic(self._num_devices)
        super(GridModel, self).__init__(mld_numeric=mld_numeric, mld_symbolic=mld_symbolic,
                                        mld_callable=mld_callable, param_struct=param_struct)

    @property
    def num_devices(self):
        return self._num_devices

    @num_devices.setter
    def num_devices(self, num_devices):
        if num_devices != self._num_devices:
            self.update_mld(mld_symbolic=self.get_grid_mld_symbolic(int(num_devices)))
            self._num_devices = int(num_devices)

    @staticmethod
    def get_grid_mld_symbolic(num_devices):
        def build():
            import sympy as sp
            lo, hi, eps = sp.symbols("P_g_min P_g_max eps")
            return MldModel(dict(D4=np.ones((1, num_devices)),
                                 F2=sp.Matrix([-lo, -(hi + eps), -hi, lo, -lo, hi]),
                                 F3=sp.Matrix([0, 0, 1, -1, 1, -1]),
                                 f5=sp.Matrix([-lo, -eps, 0, 0, -lo, hi]),
                                 G=sp.Matrix([-1, 1, 0, 0, -1, 1])), ts=0)
        return _cached(("grid", int(num_devices)), build)


class PvModel(PvMldSystemModel):
    """PV plant (reference: micro_grid_models.py:175-203): y = -P_pv_max P_pv_units omega."""

    def __init__(self, param_struct=None, mld_numeric=None, mld_callable=None, mld_symbolic=None):
        param_struct = param_struct or _par.pv_param_struct
        if mld_numeric is None and mld_callable is None and mld_symbolic is None:
            mld_symbolic = self.get_pv_mld_symbolic()
        super(PvModel, self).__init__(mld_numeric=mld_numeric, mld_symbolic=mld_symbolic, mld_callable=mld_callable,
                                      param_struct=param_struct)

    @staticmethod
    def get_pv_mld_symbolic():
        def build():
            import sympy as sp
            P_pv_max, P_pv_units = sp.symbols("P_pv_max P_pv_units")
            return MldModel(dict(D4=sp.Matrix([[-P_pv_max * P_pv_units]])), ts=0)
        return _cached(("pv",), build)


class ResDemandModel(PvMldSystemModel):
    """Residential demand (reference: micro_grid_models.py:206-234): y = P_res_ave P_res_units omega."""

    def __init__(self, param_struct=None, mld_numeric=None, mld_callable=None, mld_symbolic=None):
        param_struct = param_struct or _par.res_demand_param_struct
        if mld_numeric is None and mld_callable is None and mld_symbolic is None:
            mld_symbolic = self.get_res_demand_mld_symbolic()
        super(ResDemandModel, self).__init__(mld_numeric=mld_numeric, mld_symbolic=mld_symbolic,
                                             mld_callable=mld_callable, param_struct=param_struct)

    @staticmethod
    def get_res_demand_mld_symbolic():
        def build():
            import sympy as sp
            P_res_ave, P_res_units = sp.symbols("P_res_ave P_res_units")
            return MldModel(dict(D4=sp.Matrix([[P_res_ave * P_res_units]])), ts=0)
        return _cached(("resd",), build)
