"""``MpcController``: constraint controller + objective built from the cost-atom grammar.

Mirror of the reference's controllers/mpc_controller.py:20-101.  Linear atoms (everything the reference example
uses: ``q_mu``, ``q_u``, ``q_z``) run on the exact stage-DP / branch-and-cut kernels; separable Quadratic / L22 / L1
atoms on a scalar-state MLD run on the stage-DP kernels as convex stage terms; every other combination of the
grammar -- dense matrix weights, Linf, non-linear rate atoms, non-linear atoms on a vector-state MLD -- is assembled
into the canonical mixed-integer QP (pyhybridcontrol_b200/miqp.py) and solved by the ADMM branch-and-bound kernel
(csrc/miqp_admm.cu).
"""
import numpy as np

from ..utils.structs import ParNotSet
from .components.objective_atoms import ObjectiveAtoms
from .controller_base import (PredictiveController, ControllerBuildRequiredError, ControllerSolverError)  # noqa: F401


class MpcController(PredictiveController):
    def reset_components(self, x_k=None, omega_tilde_k=None):
        super(MpcController, self).reset_components(x_k=x_k, omega_tilde_k=omega_tilde_k)
        self._with_std_objective = True
        self._sense = "minimize"

    @property
    def std_obj_atoms(self):
        return self._std_obj_atoms

    def set_std_obj_atoms(self, objective_atoms_struct=None, **kwargs):
        if self._std_obj_atoms is not None:
            self._std_obj_atoms.set(objective_atoms_struct=objective_atoms_struct, **kwargs)
        else:
            self._std_obj_atoms = ObjectiveAtoms(self.mld_info_k, self.N_p, self.N_tilde, objective_atoms_struct,
                                                 **kwargs)
        self._build_required = True

    def update_std_obj_atoms(self, objective_weights_struct=None, **kwargs):
        if self._std_obj_atoms is not None:
            self._std_obj_atoms.update_atoms(objective_weights_struct, **kwargs)
        else:
            self._std_obj_atoms = ObjectiveAtoms(self.mld_info_k, self.N_p, self.N_tilde, objective_weights_struct,
                                                 **kwargs)
        self._build_required = True

    def set_objective(self, std_objective=ParNotSet, other_objectives=ParNotSet):
        if other_objectives not in (ParNotSet, None, []):
            raise NotImplementedError("only the standard cost atoms are supported on the GPU path")
        if std_objective is not ParNotSet:
            self._with_std_objective = std_objective is None
        self._build_required = True

    def build(self, with_std_objective=True, with_std_constraints=True, sense=None, disable_soft_constraints=False):
        self._mld_evo_matrices.update()
        self.set_objective(std_objective=None if with_std_objective else 0)
        self.set_constraints(std_evo_constaints=None if with_std_constraints else [],
                             disable_soft_constraints=disable_soft_constraints)
        sense = "minimize" if sense is None else sense
        if not (sense.lower().startswith("min") or sense.lower().startswith("max")):
            raise ValueError("Problem 'sense' must be either 'minimize' or 'maximize', got '%s'." % sense)
        self._sense = sense
        self._general_path = False
        if self._with_std_objective and self._std_obj_atoms is not None:
            for atom in self._std_obj_atoms.iter_atoms():
                if self._needs_general_path(atom):
                    self._general_path = True
        self._finish_build()

    def _needs_general_path(self, atom):
        """Linear atoms run on both exact kernels; Quadratic / L22 / L1 atoms run on the stage-DP kernels when the MLD
        is in their class, with per-step diagonal weights and no rate form.  Everything else is a general MIQP."""
        if atom.atom_type == "Linear":
            return False
        if self._sense.lower().startswith("max"):
            raise NotImplementedError("cost atom %s on '%s': maximising a convex term is not a convex problem"
                                      % (atom.atom_type, atom.var_name))
        batch = self._mld_evo_matrices.batch
        if atom.atom_type == "Linf" or not batch.stage_dp_ok or atom.is_rate_atom or atom.var_name in ("v", "z", "omega"):
            return True
        if atom.weight_type == "matrix":
            W = atom.weight_N_tilde
            if np.any(W - np.diag(np.diag(W)) != 0.0):
                return True
        return False

    def _cost_terms(self, k):
        info = self.mld_info_k
        Nt = self.N_tilde
        batch = self._mld_evo_matrices.batch
        sign = -1.0 if self._sense.lower().startswith("max") else 1.0
        cost_v = np.zeros((1, info.nv * Nt))
        w_x = w_y = None
        const = 0.0
        quad = {}
        if self._with_std_objective and self._std_obj_atoms is not None:
            prev = self.variables_k_neg1 or {}
            for atom in self._std_obj_atoms.iter_atoms():
                W = atom.weight_N_tilde
                dim = atom.dim
                if atom.atom_type != "Linear":
                    # per-step diagonal weight of e^2 (Quadratic / L22: vector weights enter squared,
                    # objective_atoms.py:320-336) or of |e| (L1, :338-347)
                    sq = atom.atom_type in ("Quadratic", "L22")
                    if atom.weight_type == "vector":
                        wd = W.ravel() ** 2 if sq else np.abs(W.ravel())
                    else:
                        wd = np.diag(W) if sq else np.abs(np.diag(W))
                    if atom.var_name == "mu" and not sq:
                        cost_v[0, batch.var_index("mu")] += wd            # mu >= 0: |mu| = mu
                        continue
                    key = atom.var_name + ("2" if sq else "1")
                    quad[key] = quad.get(key, 0.0) + wd.reshape(1, Nt, dim)
                    continue
                g = W.ravel() if atom.weight_type == "vector" else W.sum(axis=0)      # w'e | sum(W e)
                if atom.is_rate_atom:
                    # sum_k g_k' (e_k - e_{k-1}) = sum_k (g_k - g_{k+1})' e_k - g_0' e_{-1}
                    g2 = g.copy()
                    g2[:-dim] -= g[dim:]
                    e_prev = np.asarray(prev.get(atom.var_name, np.zeros((dim, 1))), dtype=float).ravel()
                    if not np.all(np.isfinite(e_prev)):
                        e_prev = np.zeros(dim)
                    const -= float(g[:dim] @ e_prev)
                    g = g2
                name = atom.var_name
                if name == "v":
                    cost_v[0] += g
                elif name in ("u", "delta", "z", "mu"):
                    cost_v[0, batch.var_index(name)] += g
                elif name == "x":
                    w_x = g[None, :] if w_x is None else w_x + g[None, :]
                elif name == "y":
                    w_y = g[None, :] if w_y is None else w_y + g[None, :]
                elif name == "omega":
                    const += float(g @ self._omega_tilde_k.ravel())
        return dict(cost_v=sign * cost_v, w_x=None if w_x is None else sign * w_x,
                    w_y=None if w_y is None else sign * w_y, const=sign * const, quad=quad or None)
