"""Agent registry + controller ownership (reference: models/agents.py:19-190)."""
from ..controllers.controller_base import ControllerBase
from ..controllers.mpc_controller import MpcController
from ..utils.structs import StructDict
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


class ControlledAgent(Agent):
    def __init__(self, device_type=None, device_id=None, sim_model=None, control_model=None):
        super(ControlledAgent, self).__init__(device_type, device_id, sim_model, control_model)
        self._controllers = StructDict()

    @property
    def controllers(self):
        return self._controllers

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
