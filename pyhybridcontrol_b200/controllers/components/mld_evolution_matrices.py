"""``MldEvoMatrices``: the condensed horizon matrices of one MLD, computed by the K1 CUDA kernel.

Same three groups / same names as the reference (controllers/components/mld_evolution_matrices.py:24-38):
``state_input`` (Phi_x, Gamma_v, Gamma_omega, Gamma_5), ``output`` (L_*), ``constraint`` (H_*), each with
``*_N_tilde`` and the row-prefix ``*_N_p`` variants (:246-250).  Values are numpy arrays copied from the
device on first access; ``.device`` keeps the CUDA tensors for the solve.
"""
import numpy as np

from ... import cabi
from ...batch import BatchMpc
from ...utils.structs import StructDict

_GROUPS = (("state_input", ("Phi_x", "Gamma_v", "Gamma_omega", "Gamma_5"), "nx"),
           ("output", ("L_x", "L_v", "L_omega", "L_5"), "ny"),
           ("constraint", ("H_x", "H_v", "H_omega", "H_5"), "n_constraints"))


def controller_model(controller):
    return controller.mld_numeric_k if controller is not None else None


class MldEvoMatrices(StructDict):
    matrix_types = StructDict(state_input="state_input", output="output", constraint="constraint")

    def __init__(self, controller=None, N_p=None, N_tilde=None, mld_numeric_k=None, mld_numeric_tilde=None,
                 device="cuda"):
        super(MldEvoMatrices, self).__init__()
        if mld_numeric_tilde:
            raise NotImplementedError("time-varying mld_numeric_tilde is dead code in the reference "
                                      "(controllers/controller_base.py:175); LTI models only")
        if controller is not None:
            N_p = controller.N_p if N_p is None else N_p
            N_tilde = controller.N_tilde if N_tilde is None else N_tilde
            mld_numeric_k = controller.mld_numeric_k if mld_numeric_k is None else mld_numeric_k
        self._N_p = int(N_p)
        self._N_tilde = int(N_tilde) if N_tilde is not None else self._N_p + 1
        # a controller's model object can be REPLACED (MldSystemModel.update_param_struct / update_mld build a new
        # numeric model), so the component keeps the controller and re-reads its model on every update, like the
        # reference's process_base_args / has_updated_version (mld_evolution_matrices.py:73-87)
        self._controller = controller if mld_numeric_k is controller_model(controller) else None
        self._mld = mld_numeric_k
        self._device = device
        self._version = None
        self.update(reset=True)

    @property
    def N_p(self):
        return self._N_p

    @property
    def N_tilde(self):
        return self._N_tilde

    @property
    def mld_info_k(self):
        return self._mld.mld_info

    @property
    def batch(self):
        return self._batch

    def _current_mld(self):
        if self._controller is not None:
            mld = self._controller.mld_numeric_k
            if mld is not None:
                return mld
        return self._mld

    def update(self, reset=False):
        """Recondense only when the model changed -- a new version of the same object or a new object (reference :73-87)."""
        mld = self._current_mld()
        key = (id(mld), mld.version)
        if not reset and self._version == key:
            return
        self._mld = mld
        info = mld.mld_info
        mats = {k: mld[k] for k in cabi.MAT_NAMES if mld[k].size}
        self._batch = BatchMpc(mats, self._N_p, self._N_tilde, nu_l=info.nu_l, nmu_l=info.nmu_l, B=1,
                               device=self._device)
        evo = self._batch.build()
        self.device = evo
        for grp, names, dim_name in _GROUPS:
            g = StructDict()
            rows_per_step = info[dim_name]
            for nm in names:
                full = evo[nm][0].cpu().numpy()
                if nm == "Phi_x" and info.nx == 0:
                    full = np.zeros((0, 0))
                g[nm + "_N_tilde"] = full
                g[nm + "_N_p"] = full[:self._N_p * rows_per_step, :]
            self[grp] = g
        self._version = key

    def get_evo_matrices_N_tilde(self, N_tilde=None):
        if N_tilde is None or N_tilde == self._N_tilde:
            return self
        if N_tilde > self._N_tilde:
            raise ValueError("N_tilde:%d cannot be greater than self.N_tilde:%d" % (N_tilde, self._N_tilde))
        out = StructDict()
        info = self.mld_info_k
        for grp, names, dim_name in _GROUPS:
            g = StructDict(self[grp])
            for nm in names:
                g[nm + "_N_tilde"] = self[grp][nm + "_N_tilde"][:N_tilde * info[dim_name], :]
            out[grp] = g
        return out
