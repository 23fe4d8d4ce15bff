"""Cost-atom grammar of the reference, host side (weights are tiny; the pull-back through the condensed
matrices happens on the GPU in hmpc_linear_cost_f64).

Key = ``<w>[_<Atom>]_[d]<var>[_N_tilde|_N_p|_f]``  (reference: controllers/components/objective_atoms.py:453-496)
  * <w> lowercase -> vector weight, default atom Linear; uppercase -> matrix weight, default atom Quadratic
  * <Atom> in Linear | Quadratic | L1 | L22 | Linf
  * a leading ``d`` (not followed by ``e``) marks a rate atom on v(k) - v(k-1)
  * suffix selects the stored weight: all steps / first N_p steps / terminal step
Weights are tiled (vector, :118-137) or block-diagonally repeated (matrix, :185-206); all-zero weights delete
the atom (:508, :519-520).
"""
import re

import numpy as np

from ...utils.structs import atleast_2d_col

VAR_NAMES = ("x", "u", "delta", "z", "omega", "y", "mu", "v")
_ATOM_RE = re.compile(r"(Linear)|(Quadratic)|([L](1|(22)|(inf)))")
_RATE_RE = re.compile(r"[dD][^e]")
_POSTFIX = ("N_p", "N_tilde", "f", "")


class ObjectiveAtom(object):
    def __init__(self, var_name, atom_type, weight_type, is_rate_atom, dim, N_p, N_tilde):
        self.var_name, self.atom_type, self.weight_type = var_name, atom_type, weight_type
        self.is_rate_atom, self.dim, self.N_p, self.N_tilde = is_rate_atom, dim, N_p, N_tilde
        shape = (dim * N_tilde, 1) if weight_type == "vector" else (dim * N_tilde, dim * N_tilde)
        self.weight_N_tilde = np.zeros(shape)

    @property
    def weight_N_p(self):
        k = self.dim * self.N_p
        return self.weight_N_tilde[:k, :1] if self.weight_type == "vector" else self.weight_N_tilde[:k, :k]

    @property
    def weight_f(self):
        k = self.dim
        return self.weight_N_tilde[-k:, :1] if self.weight_type == "vector" else self.weight_N_tilde[-k:, -k:]

    def is_zero(self):
        return bool(np.all(np.isclose(self.weight_N_tilde, 0.0)))

    def set_weight(self, value, post_fix):
        dim, Nt, N_p = self.dim, self.N_tilde, self.N_p
        value = np.asarray(atleast_2d_col(value), dtype=np.float64)
        length, lname = {"N_tilde": (Nt, "N_tilde"), "N_p": (N_p, "N_p"), "f": (1, "1")}[post_fix]
        if post_fix == "N_p" and N_p > Nt:
            raise ValueError("Cannot set weight_N_p if N_tilde < N_p")
        W = self.weight_N_tilde
        if self.weight_type == "vector":
            if value.shape[1] != 1:
                raise ValueError("Column dim of vector weight for opt_var: '%s', must be 1." % self.var_name)
            if post_fix == "f":
                if value.shape[0] != dim:
                    raise ValueError("Row dim of vector terminal weight for opt_var: '%s' must be in {%d}"
                                     % (self.var_name, dim))
                W[-dim:] = value
            else:
                if value.shape[0] == dim:
                    value = np.tile(value, (length, 1))
                elif value.shape[0] != dim * length:
                    raise ValueError("Row dim of vector weight for opt_var: '%s', must be in {%d, %d*%s}"
                                     % (self.var_name, dim, dim, lname))
                W[:dim * length] = value
        else:
            if value.shape[0] != value.shape[1]:
                raise ValueError("matrix weight for opt_var: '%s', must be square. Currently has shape: %s"
                                 % (self.var_name, value.shape))
            if post_fix == "f":
                if value.shape[0] != dim:
                    raise ValueError("Row dim of matrix terminal weight for opt_var: '%s' must be in {%d}"
                                     % (self.var_name, dim))
                W[-dim:, -dim:] = value
            else:
                if value.shape[0] == dim:
                    full = np.zeros((dim * length, dim * length))
                    for k in range(length):
                        full[k * dim:(k + 1) * dim, k * dim:(k + 1) * dim] = value
                    value = full
                elif value.shape[0] != dim * length:
                    raise ValueError("Row dim of matrix weight for opt_var: '%s', must be in {%d, %d*%s}"
                                     % (self.var_name, dim, dim, lname))
                W[:dim * length, :dim * length] = value


def parse_atom_key(key):
    info = key.split("_")
    weight_type = "vector" if "".join(info[0:1]).islower() else "matrix"
    atom_type = "".join(info[1:2]).capitalize()
    if not _ATOM_RE.search(atom_type):
        atom_type = "Linear" if weight_type == "vector" else "Quadratic"
        var_name = "".join(info[1:2]).lower()
        post_fix = "_".join(info[2:])
    else:
        var_name = "".join(info[2:3]).lower()
        post_fix = "_".join(info[3:])
    is_rate = False
    if _RATE_RE.search(var_name):
        var_name = var_name[1:]
        is_rate = True
    if var_name not in VAR_NAMES or post_fix not in _POSTFIX:
        raise ValueError("weight_name: '%s' is not valid. Must be of the form:\n  \"lower/upper[_Linear|_Quadratic|"
                         "_L1|_L22|_Linf]_[d]var_name[_N_tilde|_N_p|_f]\"" % key)
    return weight_type, atom_type, var_name, is_rate, post_fix


class ObjectiveAtoms(dict):
    """var_name -> {atom_name -> ObjectiveAtom}."""

    def __init__(self, mld_info, N_p, N_tilde, objective_atoms_struct=None, **kwargs):
        super(ObjectiveAtoms, self).__init__()
        self._info, self.N_p, self.N_tilde = mld_info, N_p, N_tilde
        self.update_atoms(objective_atoms_struct, **kwargs)

    def set(self, objective_atoms_struct=None, **kwargs):
        self.clear()
        self.update_atoms(objective_atoms_struct, **kwargs)

    def update_atoms(self, objective_atoms_struct=None, **kwargs):
        items = dict(objective_atoms_struct or {})
        items.update(kwargs)
        for key, value in items.items():
            weight_type, atom_type, var_name, is_rate, post_fix = parse_atom_key(key)
            if value is None:
                continue
            dim = self._info.get_var_dim(var_name)
            value = np.asarray(atleast_2d_col(value), dtype=np.float64)
            if not post_fix:
                post_fix = "N_tilde" if value.shape[0] in (dim, dim * self.N_tilde) else "N_p"
            atom_name = "_".join([atom_type, weight_type]) + ("_d" if is_rate else "")
            atoms = self.setdefault(var_name, {})
            atom = atoms.get(atom_name)
            if atom is None:
                if np.all(np.isclose(value, 0.0)):
                    continue
                atom = ObjectiveAtom(var_name, atom_type, weight_type, is_rate, dim, self.N_p, self.N_tilde)
                atoms[atom_name] = atom
            atom.set_weight(value, post_fix)
            if atom.is_zero():
                del atoms[atom_name]

    def iter_atoms(self):
        for var_name, atoms in self.items():
            for atom in atoms.values():
                if self._info.get_var_dim(var_name):
                    yield atom
