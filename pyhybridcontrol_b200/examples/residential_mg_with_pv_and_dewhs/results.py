"""Result format of the example: closed-loop logs of the fleet -> the DataFrame the reference's analysis scripts read.

Pure formatting, no arithmetic.  The reference keeps, per controller of every device, an
``MldSimLog`` of column vectors per step k (controllers/controller_base.py:58-146, filled by ``sim_step_k`` :229-253
and, for the grid, micro_grid_agents.py:746-756); ``get_concat_log`` turns it into a frame indexed by k with the
column MultiIndex (var_names, var_index), every agent prepends (device_type, device_id, controller)
(micro_grid_agents.py:142-150) and the grid agent concatenates itself and its devices, ordered by (type, id)
(:522-525, :551-556).  The campaign script re-indexes by time and pickles the frame under a name built from the run's
parameters (micro_grid_control_simulation.py:246-264).  This module reproduces that layout column for column from
the arrays ``DewhFleet.closed_loop`` returns, vectorised over the agents.

Variables of dimension zero have no columns (``get_concat_log`` drops empty sequences), booleans come out as floats
when any entry of the log is NaN-padded and as floats here throughout (the reference's ``cons`` column is 0/1 once
concatenated with the float columns' NaN padding).

Known difference: the simulated slack ``mu`` of a DEWH is not unique in the reference (any feasible value from its
auxiliary feasibility solve, models/mld_model.py:701-766); here it is the minimal one, max(0, violation).
"""
import os

import numpy as np

LEVELS = ("device_type", "device_id", "controller", "var_names", "var_index")

# order of LSimStruct_k (models/mld_model.py:644-645) and of VariablesStruct_k (controllers/components/variables.py:19-20)
_SIM_ORDER = ("x_k1", "x", "u", "delta", "z", "omega", "y", "mu", "v", "cons")
_HAT_ORDER = ("x", "u", "delta", "z", "omega", "y", "mu", "v")
_TIME_VARS = ("time_solve_overall", "time_in_solver")


def _as3(a, steps):
    """[steps], [steps, n] or [steps, n, dim] -> [steps, n, dim] float"""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None, None]
    elif a.ndim == 2:
        a = a[:, :, None]
    if a.shape[0] != steps:
        raise ValueError("expected %d steps, got an array of shape %s" % (steps, a.shape))
    return a


def _blocks_to_columns(device_type, device_ids, controller, blocks, steps):
    """blocks: list of (var_name, [steps, n_dev, dim]); -> values [steps, n_dev * sum(dim)] (device-major) + tuples"""
    blocks = [(name, _as3(arr, steps)) for name, arr in blocks if arr is not None]
    blocks = [(name, arr) for name, arr in blocks if arr.shape[2]]
    n_dev = len(device_ids)
    per_dev = np.concatenate([np.broadcast_to(arr, (steps, n_dev, arr.shape[2])) for _, arr in blocks], axis=2)
    names = [(name, i) for name, arr in blocks for i in range(arr.shape[2])]
    cols = [(device_type, dev_id, controller, name, i) for dev_id in device_ids for name, i in names]
    return per_dev.reshape(steps, -1), cols


def dewh_log_blocks(log, params, controller):
    """Columns of one controller of every DEWH from a ``closed_loop`` log (numpy arrays).

    Simulation part: LSimStruct_k of the re-parametrised model (micro_grid_agents.py:389-408); ``x`` is the clamped
    temperature the step used (:398-399).  ``*_hat`` part: the controller's own first-step values -- for an MPC
    controller its VariablesStruct_k (controller_base.py:243-245), for the thermostat the simulated step it keeps as
    ``variables_k`` (theromstat_control.py:62), which also carries ``cons``."""
    T = np.asarray(log["T"], dtype=np.float64)
    steps, B = T.shape[0] - 1, T.shape[1]
    T_w = np.array([p["T_w"] for p in params])
    T_max = np.array([p["T_h_max"] for p in params])
    T_min = np.array([p["T_h_min"] for p in params])
    x = np.where(T[:-1] <= T_w, T_w + 0.1, T[:-1])
    u = np.asarray(log["u"], dtype=np.float64)
    mu = np.stack([np.maximum(0.0, x - T_max), np.maximum(0.0, T_min - x)], axis=2)
    v = np.concatenate([u[:, :, None], mu], axis=2)
    cons = np.asarray(log["cons"], dtype=np.float64)
    omega = np.asarray(log["omega"], dtype=np.float64)
    sim = dict(x_k1=T[1:], x=x, u=u, omega=omega, y=x, mu=mu, v=v, cons=cons)
    blocks = [(n, sim[n]) for n in _SIM_ORDER if n in sim]
    if controller == "thermo":
        hat = dict(sim, x=T[:-1], y=T[:-1])
        hat["omega"] = np.asarray(log["omega_hat"], dtype=np.float64)
        blocks += [(n + "_hat", hat[n]) for n in _SIM_ORDER[1:] if n in hat]
    else:
        mu_hat = np.asarray(log["mu_hat"], dtype=np.float64)
        hat = dict(x=T[:-1], u=u, omega=np.asarray(log["omega_hat"], dtype=np.float64), y=T[:-1], mu=mu_hat,
                   v=np.concatenate([u[:, :, None], mu_hat], axis=2))
        blocks += [(n + "_hat", hat[n]) for n in _HAT_ORDER if n in hat]
    # one batched solve serves every agent: its wall / device time is reported for each of them
    t = np.asarray(log.get("solve_ms", np.zeros(steps)), dtype=np.float64) * 1e-3
    t_all = np.asarray(log.get("step_ms", log.get("solve_ms", np.zeros(steps))), dtype=np.float64) * 1e-3
    blocks += [("time_solve_overall", np.broadcast_to(t_all[:, None], (steps, B))),
               ("time_in_solver", np.broadcast_to(t[:, None], (steps, B)))]
    return blocks


def source_log_blocks(omega, omega_hat, gain, is_mpc=True, times=None):
    """PV / residential-demand device (nx = nu = 0, y = gain * omega; micro_grid_models.py:176-240):
    PV gain = -P_pv_max * P_pv_units, demand gain = P_res_ave * P_res_units.

    The ``*_hat`` columns are there for every controller type: the ``NoController`` these devices get next to a
    thermostat (micro_grid_control_simulation.py:175-177) is a plain ConstraintSolvedController
    (controllers/no_controller.py) and logs its forecast like an MPC controller does -- seen in the frame of the
    reference's real loop (tests/golden/microgrid_loop.npz).  ``is_mpc=False`` with ``omega_hat=None`` leaves them
    out."""
    omega = np.asarray(omega, dtype=np.float64).reshape(-1)
    steps = omega.size
    blocks = [("omega", omega), ("y", gain * omega)]
    if is_mpc or omega_hat is not None:
        omega_hat = np.asarray(omega_hat, dtype=np.float64).reshape(-1)
        blocks += [("omega_hat", omega_hat), ("y_hat", gain * omega_hat)]
    times = np.zeros((steps, 2)) if times is None else np.asarray(times, dtype=np.float64)
    blocks += [("time_solve_overall", times[:, 0]), ("time_in_solver", times[:, 1])]
    return blocks


def grid_log_blocks(grid, is_mpc=True, times=None):
    """Grid agent (micro_grid_models.py:137-172: ndelta = nz = ny = 1, nomega = number of devices, 6 rows) from the
    arrays of ``DewhFleet.grid_log``: omega [steps, n_dev] device powers, y, delta, z, cons [steps, 6], the ``*_hat``
    counterparts from the planned first step, and price [steps]; adds p_imp / p_exp / cost
    (micro_grid_agents.py:746-756)."""
    y, delta, z = (np.asarray(grid[k], dtype=np.float64).reshape(-1) for k in ("y", "delta", "z"))
    steps = y.size
    v = np.stack([delta, z], axis=1)
    blocks = [("delta", delta), ("z", z), ("omega", np.asarray(grid["omega"], dtype=np.float64)[:, None, :]), ("y", y),
              ("v", v[:, None, :]), ("cons", np.asarray(grid["cons"], dtype=np.float64)[:, None, :])]
    if is_mpc or "y_hat" in grid:                          # a NoController grid logs its forecast too (see above)
        yh, dh, zh = (np.asarray(grid[k + "_hat"], dtype=np.float64).reshape(-1) for k in ("y", "delta", "z"))
        blocks += [("delta_hat", dh), ("z_hat", zh),
                   ("omega_hat", np.asarray(grid["omega_hat"], dtype=np.float64)[:, None, :]), ("y_hat", yh),
                   ("v_hat", np.stack([dh, zh], axis=1)[:, None, :])]
    times = np.zeros((steps, 2)) if times is None else np.asarray(times, dtype=np.float64)
    blocks += [("time_solve_overall", times[:, 0]), ("time_in_solver", times[:, 1])]
    price = np.asarray(grid["price"], dtype=np.float64).reshape(-1)
    blocks += [("p_imp", z), ("p_exp", y - z), ("cost", z * price)]
    return blocks


def grid_sim_dataframe(device_blocks, steps, time_0=None, freq="15min"):
    """device_blocks: list of (device_type, device_ids, controller, blocks) in any order -> the frame of
    ``GridAgentMpc.grid_sim_dataframe`` (micro_grid_agents.py:522-525): the grid first, then the devices ordered by
    (type, id), for each device its controllers in the order given, columns named LEVELS.  With ``time_0`` the index
    is the time range the campaign script sets (micro_grid_control_simulation.py:246-247), else k = 0..steps-1."""
    import pandas as pd
    per_device = {}
    for device_type, device_ids, controller, blocks in device_blocks:
        vals, cols = _blocks_to_columns(device_type, list(device_ids), controller, blocks, steps)
        width = len(cols) // len(device_ids)
        for i, dev_id in enumerate(device_ids):
            per_device.setdefault((device_type, dev_id), []).append(
                (vals[:, i * width:(i + 1) * width], cols[i * width:(i + 1) * width]))
    order = sorted((key for key in per_device if key[0] == "grid"), key=lambda t: t[1])
    order += sorted((key for key in per_device if key[0] != "grid"), key=lambda t: (t[0], t[1]))
    values = np.concatenate([v for key in order for v, _ in per_device[key]], axis=1)
    columns = [c for key in order for _, cs in per_device[key] for c in cs]
    df = pd.DataFrame(values, columns=pd.MultiIndex.from_tuples(columns, names=LEVELS))
    if time_0 is not None:
        df.index = pd.date_range(start=time_0, periods=steps, freq=freq)
    else:
        df.index.name = "k"
    return df


def sim_out_path(base_dir, N_p, soft_top_mult, soft_bot_mult, num_scenarios, N_sb_reduced, N_h, T_max, T_min,
                 save_text_postfix=""):
    """File name of a campaign result (micro_grid_control_simulation.py:249-262)."""
    if save_text_postfix and not save_text_postfix.startswith("_"):
        save_text_postfix = "_" + save_text_postfix
    name = ("sim_Np_%d_st_%d_sb_%d_Ns_%d_Nsr_%d_Nh_%d_Tmax_%d_Tmin_%d%s.sim_out"
            % (N_p, int(soft_top_mult), int(soft_bot_mult), num_scenarios, N_sb_reduced, N_h, int(T_max), int(T_min),
               save_text_postfix))
    return os.path.realpath(os.path.join(base_dir, "sim_out", name))


def save_sim_out(df, path):
    """``df_sim.to_pickle(save_path)`` with the directory created first (:254-264)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    df.to_pickle(path)
    return path
