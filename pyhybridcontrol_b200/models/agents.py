"""Agent registry + controller ownership (reference: models/agents.py:19-190)."""
from ..controllers.controller_base import ControllerBase
from ..controllers.mpc_controller import MpcController
from ..utils.structs import ParNotSet, StructDict
from .mld_model import MldSystemModel


class Agent(object):
    _device_type_id_struct = StructDict()

    def __init__(self, device_type=None, device_id=None, sim_model=None, control_model=None):
        self._device_type = device_type if device_type is not None else "not_specified"
        ids = self._device_type_id_struct.setdefault(self._device_type, set())
        if device_id is None:
            device_id = max(ids) + 1 if ids else 1
        if device_id in ids:
            raise ValueError("Agent with type:%s and device_id:%s already exists." % (self._device_type, device_id))
        ids.add(device_id)
        self._device_id = device_id
        self._sim_model = sim_model if sim_model is not None else MldSystemModel()
        self._control_model = control_model

    @classmethod
    def delete_all_devices(cls):
        cls._device_type_id_struct.clear()

    def update_device_data(self, device_type=None, device_id=None):
        """Re-register under another type / id (models/agents.py:39-59): the new pair must be free."""
        new_type = device_type if device_type is not None else self._device_type or "not_specified"
        ids = self._device_type_id_struct.setdefault(new_type, set())
        new_id = device_id if device_id is not None else self._device_id
        same = new_type == self._device_type and new_id == self._device_id
        if new_id is None:
            new_id = max(ids) + 1 if ids else 1
        elif new_id in ids and not same:
            raise ValueError("Agent with type:'%s' and device_id:'%s' already exists" % (new_type, new_id))
        self._device_type_id_struct.get(self._device_type, set()).discard(self._device_id)
        ids.add(new_id)
        self._device_type, self._device_id = new_type, new_id

    def update_models(self, sim_model=ParNotSet, control_model=ParNotSet):
        """Swap the simulation / control model (models/agents.py:61-69); owners of controllers reset them."""
        if sim_model is not ParNotSet:
            self._sim_model = sim_model if sim_model is not None else MldSystemModel()
        if control_model is not ParNotSet:
            self._control_model = control_model

    def __del__(self):
        try:
            self._device_type_id_struct[self._device_type].discard(self._device_id)
        except Exception:
            pass

    @property
    def device_type(self):
        return self._device_type

    @property
    def device_id(self):
        return self._device_id

    @property
    def sim_model(self):
        return self._sim_model

    @property
    def control_model(self):
        return self._control_model if self._control_model is not None else self._sim_model

    @property
    def mld_numeric(self):
        return self.control_model.mld_numeric

    @property
    def mld_info(self):
        return self.mld_numeric.mld_info

    @property
    def mld_numeric_tilde(self):
        return None                       # (time-varying models along the horizon: dead code in the reference too)


class ControlledAgent(Agent):
    def __init__(self, device_type=None, device_id=None, sim_model=None, control_model=None):
        super(ControlledAgent, self).__init__(device_type, device_id, sim_model, control_model)
        self._controllers = StructDict()

    @property
    def controllers(self):
        return self._controllers

    def update_models(self, sim_model=ParNotSet, control_model=ParNotSet):
        """A model swap invalidates every controller's condensed matrices and atoms (models/agents.py:124-128)."""
        super(ControlledAgent, self).update_models(sim_model=sim_model, control_model=control_model)
        for controller in self._controllers.values():
            controller.reset_components()

    def add_controller(self, name, controller_type, x_k=None, omega_tilde_k=None, N_p=None, N_tilde=None, **kwargs):
        if not (isinstance(controller_type, type) and issubclass(controller_type, ControllerBase)):
            raise TypeError("controller_type must be a subclass of ControllerBase")
        self._controllers[name] = controller_type(agent=self, N_p=N_p, N_tilde=N_tilde, x_k=x_k,
                                                  omega_tilde_k=omega_tilde_k, **kwargs)
        return self._controllers[name]

    def delete_controller(self, name):
        del self._controllers[name]

    def delete_all_controllers(self):
        self._controllers.clear()


class MpcAgent(ControlledAgent):
    """Single-MPC convenience wrapper (reference: models/agents.py:145-190)."""

    def __init__(self, device_type=None, device_id=None, sim_model=None, control_model=None, N_p=None, N_tilde=None,
                 **kwargs):
        super(MpcAgent, self).__init__(device_type, device_id, sim_model, control_model)
        self.add_controller("mpc", MpcController, N_p=N_p, N_tilde=N_tilde, **kwargs)

    @property
    def mpc_controller(self) -> MpcController:
        return self._controllers["mpc"]

    @property
    def N_p(self):
        return self.mpc_controller.N_p

    @property
    def N_tilde(self):
        return self.mpc_controller.N_tilde

    @property
    def x_k(self):
        return self.mpc_controller.x_k

    @x_k.setter
    def x_k(self, value):
        self.mpc_controller.x_k = value

    @property
    def omega_tilde_k(self):
        return self.mpc_controller.omega_tilde_k

    @omega_tilde_k.setter
    def omega_tilde_k(self, value):
        self.mpc_controller.omega_tilde_k = value

    # the reference's name for the forecast the controller works with (models/agents.py:185-190)
    omega_tilde_k_hat = omega_tilde_k

    def update_horizons(self, N_p=ParNotSet, N_tilde=ParNotSet):
        """models/agents.py:159-162: N_tilde follows N_p unless given."""
        N_p = N_p if N_p is not ParNotSet else self.N_p or 0
        N_tilde = N_tilde if N_tilde is not ParNotSet else N_p + 1
        self.mpc_controller.update_horizons(N_p=N_p, N_tilde=N_tilde)
