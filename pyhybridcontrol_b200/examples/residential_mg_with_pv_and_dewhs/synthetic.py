"""Seeded synthetic DEWH workload of the example's shape.

The reference's two DHW-demand pickles are missing from its checkout (.MISSING_LARGE_BLOBS), so demand is
synthetic: zero-inflated log-normal draws scaled to 200 L/day (file names say ``200Lpd_mean``), expressed in
L/s like examples/.../micro_grid_control_simulation.py:37-40.  Prices are CONTINUOUS (log-normal around the
time-of-use levels, converted like :89-90) so that optima are unique and decision parity is meaningful.
"""
import numpy as np

from .parameters import dewh_param_struct, TOU_LEVELS

STEPS_PER_DAY = 96


def price_profile(num_steps, seed=0, ts=900.0):
    """-> (num_steps,) energy price per W per control step [currency/W/step], continuous."""
    rng = np.random.default_rng([seed, 0x5052])
    k = np.arange(num_steps) % STEPS_PER_DAY
    hour = k / 4.0
    level = np.full(num_steps, TOU_LEVELS["low_off_peak"])
    stnd = ((hour >= 6) & (hour < 7)) | ((hour >= 10) & (hour < 18)) | ((hour >= 20) & (hour < 22))
    peak = ((hour >= 7) & (hour < 10)) | ((hour >= 18) & (hour < 20))
    level[stnd] = TOU_LEVELS["low_stnd"]
    level[peak] = TOU_LEVELS["low_peak"]
    level = level * np.exp(0.15 * rng.standard_normal(num_steps))
    return level / 3600.0 / 100.0 / 1000.0 * ts


def dhw_demand_profile(num_steps, seed, ts=900.0, litres_per_day=200.0, p_draw=0.3):
    """-> (num_steps,) hot-water draw in L/s (zero-inflated log-normal, mean ``litres_per_day``)."""
    rng = np.random.default_rng([seed, 0x4448])
    draw = rng.random(num_steps) < p_draw
    mag = np.exp(0.8 * rng.standard_normal(num_steps))
    mean_per_slot = litres_per_day / STEPS_PER_DAY
    litres = draw * mag * (mean_per_slot / (p_draw * np.exp(0.32)))
    return litres / ts


def dewh_agent_params(agent_id, jitter=0.10, T_h_max=None):
    """Per-agent parameter dict, jittered +-``jitter`` around parameters.py (seed = agent id)."""
    rng = np.random.default_rng([int(agent_id), 0x4457])
    p = dict(dewh_param_struct)
    for key in ("m_h", "U_h", "P_h_Nom"):
        p[key] = p[key] * (1.0 + jitter * rng.uniform(-1.0, 1.0))
    p["T_h_min"] = 50.0 * (1.0 + 0.2 * jitter * rng.uniform(-1.0, 1.0))
    tmax = (65.0 if rng.random() < 0.5 else 80.0) if T_h_max is None else T_h_max
    p["T_h_max"] = tmax * (1.0 + 0.2 * jitter * rng.uniform(-1.0, 1.0))
    return p


def dewh_initial_state(agent_id):
    """x0 in {55..64} C (cf. micro_grid_control_simulation.py:68-70)."""
    rng = np.random.default_rng([int(agent_id), 0x5830])
    return float(rng.integers(55, 65))


def dewh_scalars(p, const_heat=True, T_h=None, D_h=None):
    """Discretised scalar DEWH model (A, B1, B4, b5); reference: examples/.../micro_grid_models.py:27-63."""
    p1 = p["U_h"] * p["A_h"]
    p2 = p["m_h"] * p["C_w"]
    if const_heat:
        a_c = -p1 / p2
        b4_c = p["C_w"] * (p["T_w"] - p["T_h_Nom"]) / p2
    else:
        T_h = p["T_h"] if T_h is None else T_h
        D_h = p["D_h"] if D_h is None else D_h
        r = (p["T_h_Nom"] - p["T_w"]) / (T_h - p["T_w"])
        a_c = -((D_h * p["C_w"] * r) + p1) / p2
        b4_c = p["C_w"] * p["T_w"] * r / p2
    A = np.exp(a_c * p["ts"])
    em = (A - 1.0) / a_c
    return A, em * p["P_h_Nom"] / p2, em * b4_c, em * p1 * p["T_inf"] / p2


def dewh_batch(B, N_p, seed=0, k0=0, soft_top_mult=10.0, soft_bot_mult=1.0, first_agent=0):
    """Everything a batch of ``B`` DEWH control steps needs, as float64 numpy arrays.

    Returns a dict: mats (name -> [B, r, c]), x0 [B,1], omega [B, Nt], q_u [B, Nt], q_mu [B, 2], params list.
    Cost follows the example: q_u = price * P_h_Nom, q_mu = [10, 1] * sum(q_u)
    (micro_grid_control_simulation.py:193-198).
    """
    Nt = N_p + 1
    price = price_profile(k0 + Nt, seed=seed)[k0:k0 + Nt]
    mats = {k: np.zeros((B,) + s) for k, s in (("A", (1, 1)), ("B1", (1, 1)), ("B4", (1, 1)), ("b5", (1, 1)),
                                                 ("E", (2, 1)), ("F1", (2, 1)), ("Psi", (2, 2)), ("f5", (2, 1)))}
    x0 = np.zeros((B, 1))
    omega = np.zeros((B, Nt))
    q_u = np.zeros((B, Nt))
    q_mu = np.zeros((B, 2))
    params = []
    for b in range(B):
        aid = first_agent + b
        p = dewh_agent_params(aid)
        params.append(p)
        A, B1, B4, b5 = dewh_scalars(p, const_heat=True)
        mats["A"][b, 0, 0], mats["B1"][b, 0, 0], mats["B4"][b, 0, 0], mats["b5"][b, 0, 0] = A, B1, B4, b5
        mats["E"][b] = [[1.0], [-1.0]]
        mats["Psi"][b] = [[-1.0, 0.0], [0.0, -1.0]]
        mats["f5"][b] = [[p["T_h_max"]], [-p["T_h_min"]]]
        x0[b, 0] = dewh_initial_state(aid)
        omega[b] = dhw_demand_profile(k0 + Nt, seed=aid)[k0:k0 + Nt]
        # per-agent multiplicative price jitter keeps every agent's optimum unique as well
        rng = np.random.default_rng([aid, 0x5155])
        q_u[b] = price * p["P_h_Nom"] * np.exp(0.02 * rng.standard_normal(Nt))
        q_mu[b] = [soft_top_mult * q_u[b].sum(), soft_bot_mult * q_u[b].sum()]
    return dict(mats=mats, x0=x0, omega=omega, q_u=q_u, q_mu=q_mu, params=params, N_p=N_p, Nt=Nt, B=B)
