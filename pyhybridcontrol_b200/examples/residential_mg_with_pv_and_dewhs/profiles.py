"""Input side of the example's control loop: disturbance / price profiles -> the forecast windows, actual draws,
scenario sets and min / max windows every control instant needs, for a whole fleet at once.

Reference (per device, per instant, pandas frames):
  * ``MicroGridAgentBase.set_omega_profile``          examples/.../modelling/micro_grid_agents.py:190-204
  * ``get_omega_tilde_k_act``   (window shifted by ``forecast_lag``: what really happens)      :236-263
  * ``get_omega_tilde_k_hat``   (window at k: yesterday's draw is today's forecast; a deterministic controller gets
                                 the actual one)                                                    :265-298
  * ``set_omega_scenarios`` / ``get_omega_tilde_scenario`` (a day-per-column table, random columns, windows run on
                                 into the following columns)                                        :156-188, 206-232
  * ``GridAgentMpc.set_price_profile`` / ``get_price_tilde_k``                                     :563-606
  * the script's helpers ``get_actual_omega_dewh_profiles``, ``get_dewh_random_initial_state``,
    ``get_min_max_dhw_scenario``      examples/.../micro_grid_control_simulation.py:56-83

Everything here is index arithmetic on host arrays (no floating-point work); the results are the ``demand``,
``demand_actual``, ``scenarios``, ``demand_minmax`` and ``price`` arguments of ``DewhFleet.closed_loop`` / ``campaign``.
Layout: profiles are ``[B, n, nomega]`` (device, step, disturbance); a window is ``[B, N_tilde * nomega]`` stacked by
step like the reference's ``omega_tilde`` column vectors.
"""
import re

import numpy as np

_UNITS = dict(D=86400.0, d=86400.0, H=3600.0, h=3600.0, T=60.0, min=60.0, S=1.0, s=1.0)


def lag_steps(forecast_lag, ts):
    """``int(pd.Timedelta(forecast_lag) / pd.Timedelta(seconds=ts))`` (micro_grid_agents.py:238) for lags written as
    '<number><D|H|T|min|S>' ('1D' in the reference), a number of seconds, or anything with ``total_seconds()``."""
    if hasattr(forecast_lag, "total_seconds"):
        seconds = forecast_lag.total_seconds()
    elif isinstance(forecast_lag, str):
        m = re.fullmatch(r"\s*([0-9.]*)\s*([A-Za-z]+)\s*", forecast_lag)
        if not m or m.group(2) not in _UNITS:
            raise ValueError("cannot read forecast_lag %r" % (forecast_lag,))
        seconds = float(m.group(1) or 1.0) * _UNITS[m.group(2)]
    else:
        seconds = float(forecast_lag)
    return int(seconds / float(ts))


def _as_profiles(profile, nomega, what):
    """[n], [n, nomega] (one device) or [B, n, nomega] -> [B, n, nomega]"""
    a = np.asarray(profile, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("%s must be [n], [n, nomega] or [B, n, nomega]" % what)
    if a.shape[2] != nomega:
        raise ValueError("%s must have column dimension of 'nomega':%d not %d." % (what, nomega, a.shape[2]))
    return a


class OmegaProfiles(object):
    """The disturbance profiles of B devices and the windows the loop cuts out of them."""

    def __init__(self, omega_profile, ts, nomega=1, forecast_lag="1D"):
        self.nomega = int(nomega)
        self.values = _as_profiles(omega_profile, self.nomega, "Omega profile")
        self.ts = float(ts)
        self.lag = lag_steps(forecast_lag, ts)

    @property
    def B(self):
        return self.values.shape[0]

    def _window(self, start, N_tilde):
        w = self.values[:, start:start + N_tilde, :]
        return w.reshape(self.B, -1)            # short at the end of the profile, like the reference's slice

    def omega_tilde_k_act(self, k, N_tilde):
        """what will really be drawn over the horizon: rows ``lag + k ... lag + k + N_tilde`` (:237-240)"""
        return self._window(self.lag + int(k), int(N_tilde))

    def omega_tilde_k_hat(self, k, N_tilde, deterministic=False):
        """the controller's forecast: rows ``k ... k + N_tilde``, or the actual window for a deterministic
        (perfect-forecast) controller (:267-296)"""
        return self.omega_tilde_k_act(k, N_tilde) if deterministic else self._window(int(k), int(N_tilde))

    def omega_k_act(self, k):
        """the draw applied to the simulation model at instant k = first block of the actual window"""
        return self.values[:, self.lag + int(k), :]

    def closed_loop_arrays(self, sim_steps, N_tilde):
        """``demand`` (forecast source) and ``demand_actual`` [B, sim_steps + N_tilde] of ``DewhFleet.closed_loop``
        for a scalar disturbance: closed_loop cuts ``demand[:, k:k+Nt]`` as the forecast and applies
        ``demand_actual[:, k]``, i.e. exactly ``omega_tilde_k_hat(k)`` and ``omega_k_act(k)``."""
        if self.nomega != 1:
            raise ValueError("closed_loop takes one disturbance per device")
        need = int(sim_steps) + int(N_tilde)
        if self.values.shape[1] < self.lag + need:
            raise ValueError("profile too short: %d rows, %d needed" % (self.values.shape[1], self.lag + need))
        flat = self.values[:, :, 0]
        return dict(demand=np.ascontiguousarray(flat[:, :need]),
                    demand_actual=np.ascontiguousarray(flat[:, self.lag:self.lag + need]))


class OmegaScenarios(object):
    """Historical scenario days of one device type (:156-188): the profile is cut into days, one per column, and a
    scenario window starts at the instant's time of day in a random column and runs on into the next columns."""

    def __init__(self, omega_scenarios_profile, ts, nomega=1):
        prof = np.asarray(omega_scenarios_profile, dtype=np.float64)
        if prof.ndim == 1:
            prof = prof[:, None]
        n, m = prof.shape
        if m != int(nomega):
            raise ValueError("omega_scenarios_profile must have column dimension of 'nomega':%d not %d." % (nomega, m))
        self.nomega = m
        self.intervals_per_day = int(86400.0 // float(ts))
        self.num_scenarios = n // self.intervals_per_day
        if n % self.intervals_per_day:
            # the reference reshapes the whole stack into intervals_per_day*m rows and fails on a ragged last day
            raise ValueError("cannot reshape array of size %d into shape (%d,newaxis)"
                             % (n * m, self.intervals_per_day * m))
        # column d = day d, rows stacked by step then disturbance (row-major stack of the frame, Fortran reshape)
        self.table = np.asfortranarray(prof.reshape(self.num_scenarios, self.intervals_per_day * m).T)

    def draw_columns(self, k, N_tilde, num_scenarios=1, random_state=None):
        """the random day columns of ``get_omega_tilde_scenario`` (:206-223): same validity rule, same generator call
        (``randint(low=0, high=valid_columns, size=num_scenarios)`` on ``numpy.random`` unless a RandomState /
        Generator-like object with ``randint`` is given)"""
        span = int(N_tilde) * self.nomega
        row = (int(k) % self.intervals_per_day) * self.nomega
        limit = self.table.size - row - span - 1
        if limit <= 0 or limit < span * int(num_scenarios):
            raise ValueError("Insufficient number of scenarios to draw from.")
        valid_columns = limit // self.table.shape[0] - 1
        rs = np.random if random_state is None else random_state
        return row, rs.randint(low=0, high=valid_columns, size=int(num_scenarios))

    def omega_tilde_scenario(self, k, N_tilde, num_scenarios=1, random_state=None):
        """[N_tilde * nomega, num_scenarios] -- one device's scenario set at instant k (:206-232)"""
        row, cols = self.draw_columns(k, N_tilde, num_scenarios, random_state)
        flat = self.table.ravel(order="F")
        span = int(N_tilde) * self.nomega
        if len(cols) == 0:
            return None
        return np.stack([flat[self.table.shape[0] * c + row:self.table.shape[0] * c + row + span] for c in cols], axis=1)

    def fleet_scenarios(self, k, N_tilde, num_scenarios, B, random_state=None):
        """[B, N_tilde * nomega, num_scenarios]: every device draws its own set, in device order, as the script's
        loop does (micro_grid_control_simulation.py:199-201) -- the ``scenarios(k)`` callable of closed_loop."""
        return np.stack([self.omega_tilde_scenario(k, N_tilde, num_scenarios, random_state) for _ in range(int(B))])

    def min_max_day(self):
        """per-time-of-day minimum and maximum over the scenario days (micro_grid_control_simulation.py:110-111)"""
        return self.table.min(axis=1), self.table.max(axis=1)


class PriceProfile(object):
    """The grid's price profile and its windows (micro_grid_agents.py:563-606): like the actual disturbance, the
    price window is shifted by ``forecast_lag``."""

    def __init__(self, price_profile, ts, forecast_lag="1D"):
        a = np.asarray(price_profile, dtype=np.float64)
        self.values = a.reshape(a.shape[0], -1)
        self.lag = lag_steps(forecast_lag, ts)

    def price_tilde_k(self, k, N_tilde):
        start = self.lag + int(k)
        return self.values[start:start + int(N_tilde)].reshape(-1)

    def closed_loop_array(self, sim_steps, N_tilde):
        """``price`` [sim_steps + N_tilde] of ``DewhFleet.closed_loop`` (it cuts ``price[k:k+Nt]``)"""
        need = int(sim_steps) + int(N_tilde)
        if self.values.shape[0] < self.lag + need:
            raise ValueError("price profile too short: %d rows, %d needed" % (self.values.shape[0], self.lag + need))
        if self.values.shape[1] != 1:
            raise ValueError("closed_loop takes one price per step")
        return np.ascontiguousarray(self.values[self.lag:self.lag + need, 0])


def get_actual_omega_dewh_profiles(actual_scenarios, N_h=1, size=1):
    """device id (1..N_h) -> [rows * size, 1]: ``size`` scenario days drawn without replacement by
    ``RandomState(id**2)`` and laid end to end (micro_grid_control_simulation.py:56-65)"""
    actual_scenarios = np.asarray(getattr(actual_scenarios, "values", actual_scenarios))
    num_scen = actual_scenarios.shape[1]
    out = {}
    for i in range(1, int(N_h) + 1):
        rs = np.random.RandomState(seed=np.int32(i ** 2))
        out[i] = actual_scenarios[:, rs.choice(num_scen, size=size, replace=False)].reshape(-1, 1, order="F")
    return out


def get_dewh_random_initial_state(dev_id):
    """initial tank temperature in {55..64} from ``RandomState(dev_id**2)`` (micro_grid_control_simulation.py:68-70)"""
    return np.random.RandomState(seed=np.int32(dev_id ** 2)).randint(55, 65)


def get_min_max_dhw_scenario(k, N_tilde, min_dhw_day, max_dhw_day, steps_per_day=96):
    """[min window, max window] as columns [N_tilde, 1]: the day profiles tiled and rolled to the instant's time of day
    (micro_grid_control_simulation.py:73-83)"""
    out = []
    for day in (min_dhw_day, max_dhw_day):
        day = np.asarray(day).flatten()
        if len(day) != steps_per_day:
            raise ValueError("Invalid shape for %s" % ("min_dhw_day" if not out else "max_dhw_day"))
        out.append(day)
    pos = int(k) % steps_per_day
    mult = int(N_tilde) // steps_per_day + 1
    return [np.roll(np.tile(day, mult), -pos)[:int(N_tilde)].reshape(-1, 1) for day in out]


def min_max_closed_loop_arrays(sim_steps, N_tilde, min_dhw_day, max_dhw_day, steps_per_day=96):
    """``demand_minmax`` of ``DewhFleet.closed_loop``: two profiles [sim_steps + N_tilde] whose window ``[k:k+Nt]``
    equals ``get_min_max_dhw_scenario(k, Nt, ...)`` for every k (the day profile repeated from k = 0)."""
    need = int(sim_steps) + int(N_tilde)
    reps = need // steps_per_day + 1
    return tuple(np.tile(np.asarray(day).flatten(), reps)[:need] for day in (min_dhw_day, max_dhw_day))


# ---- time-of-use tariff (reference: examples/.../tariff_generator.py:14-123) ----------------------------------------
TARIFF_TYPES = ("low_off_peak", "low_stnd", "low_peak", "high_off_peak", "high_stnd", "high_peak")
_OFF, _STD, _PEAK = 0, 1, 2


def _hour_table():
    """[season (low, high), day kind (weekday, Saturday, Sunday), hour] -> off-peak / standard / peak"""
    tab = np.zeros((2, 3, 24), dtype=np.int64)                    # Sundays: off-peak all day (:105-110)
    hours = np.arange(24)
    # weekdays, low season (:75-85): off-peak 22-06, standard 06-07, 10-18, 20-22, peak otherwise
    wd_low = np.full(24, _PEAK)
    wd_low[(hours >= 22) | (hours < 6)] = _OFF
    wd_low[(hours == 6) | ((hours >= 10) & (hours < 18)) | ((hours >= 20) & (hours < 22))] = _STD
    # weekdays, high season (:65-74): off-peak 22-06, standard 09-17 and 19-22, peak otherwise
    wd_high = np.full(24, _PEAK)
    wd_high[(hours >= 22) | (hours < 6)] = _OFF
    wd_high[((hours >= 9) & (hours < 17)) | ((hours >= 19) & (hours < 22))] = _STD
    # Saturdays, both seasons (:87-103): off-peak 20-07 and 12-18, standard otherwise
    sat = np.full(24, _STD)
    sat[(hours >= 20) | (hours < 7) | ((hours >= 12) & (hours < 18))] = _OFF
    tab[0, 0], tab[1, 0], tab[0, 1], tab[1, 1] = wd_low, wd_high, sat, sat
    return tab


def tou_price_vector(date_time_0, n_steps, control_ts, low_off_peak=0.0, low_stnd=0.0, low_peak=0.0,
                     high_off_peak=0.0, high_stnd=0.0, high_peak=0.0):
    """[n_steps, 1] import price at ``date_time_0 + i * control_ts`` -- ``TariffGenerator(...).get_price_vector``
    (tariff_generator.py:37-47): high season = 1 June .. 31 August (:28-29, :53-55), rate by day kind and hour.  The
    example scales it by ``/3600/100/1000*ts`` from c/kWh to currency per W per step
    (micro_grid_control_simulation.py:89-90)."""
    import datetime as _dt
    if not isinstance(date_time_0, _dt.datetime):
        raise TypeError("date_time_0 must be of type 'datetime'")
    step = control_ts if isinstance(control_ts, _dt.timedelta) else _dt.timedelta(seconds=control_ts)
    t0 = np.datetime64(date_time_0, "us")
    times = t0 + np.arange(int(n_steps)) * np.timedelta64(int(round(step.total_seconds() * 1e6)), "us")
    days = times.astype("datetime64[D]")
    hour = ((times - days) // np.timedelta64(1, "h")).astype(np.int64)
    weekday = (days.astype(np.int64) + 3) % 7                       # 1970-01-01 was a Thursday; Monday = 0
    kind = np.where(weekday < 5, 0, weekday - 4)
    months = times.astype("datetime64[M]")
    month = months.astype(np.int64) % 12 + 1
    high = ((month >= 6) & (month <= 8)).astype(np.int64)            # (6, 1) <= (month, day) <= (8, 31)
    rates = np.array([[low_off_peak, low_stnd, low_peak], [high_off_peak, high_stnd, high_peak]], dtype=np.float64)
    return rates[high, _hour_table()[high, kind, hour]].reshape(-1, 1)
