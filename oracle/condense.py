"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the reference's horizon condensing
(controllers/components/mld_evolution_matrices.py).  All stacks start at step 0, so the first block of
Phi_x is I and the first block-row of every Gamma is zero (quirk list).

  x~ = Phi_x x0 + Gamma_v v~ + Gamma_omega w~ + Gamma_5
  y~ = L_x   x0 + L_v     v~ + L_omega     w~ + L_5
  H_v v~ <= H_x x0 + H_omega w~ + H_5

Pinned against the unmodified reference: tests/golden/condense_*.npz (tests/test_oracle_golden.py).
"""
import numpy as np


def a_powers(A, Nt):
    """(I, A, A^2, ...) by running left-to-right products -- mld_evolution_matrices.py:253-272."""
    nx = A.shape[0]
    P = [np.eye(nx)]
    for _ in range(Nt - 1):
        P.append(P[-1] @ A)
    return P


def _toeplitz_lower(blocks, Nt, nr, ncols):
    """Block (i, j) = blocks[i-j-1] for i > j else 0 -- _gen_input_evo_mat :467-501 + block_toeplitz
    (utils/matrix_utils.py:117-161)."""
    out = np.zeros((nr * Nt, ncols * Nt))
    for i in range(1, Nt):
        for j in range(i):
            out[i * nr:(i + 1) * nr, j * ncols:(j + 1) * ncols] = blocks[i - j - 1]
    return out


def _blkdiag(M, Nt):
    """_gen_mat_tilde_diag :504-527 (block_diag_dense, utils/matrix_utils.py:71-81)."""
    r, c = M.shape
    out = np.zeros((r * Nt, c * Nt))
    for k in range(Nt):
        out[k * r:(k + 1) * r, k * c:(k + 1) * c] = M
    return out


def condense(full, dims, Nt):
    """full/dims from oracle.mld.complete -> dict of the 12 ``*_N_tilde`` matrices."""
    nx, ny, nc, nv, nw = dims["nx"], dims["ny"], dims["nc"], dims["nv"], dims["nomega"]
    A = full["A"]
    Ap = a_powers(A, Nt)
    Bv = np.hstack([full["B1"], full["B2"], full["B3"], np.zeros((nx, dims["nmu"]))])      # :291
    Dv = np.hstack([full["D1"], full["D2"], full["D3"], np.zeros((ny, dims["nmu"]))])      # :355
    Fv = np.hstack([full["F1"], full["F2"], full["F3"], full["Psi"]])                       # :411
    out = {}
    out["Phi_x"] = np.vstack(Ap) if nx else np.zeros((0, 0))                               # :275-280
    out["Gamma_v"] = _toeplitz_lower([Ap[k] @ Bv for k in range(Nt - 1)], Nt, nx, nv)      # :283-297
    out["Gamma_omega"] = _toeplitz_lower([Ap[k] @ full["B4"] for k in range(Nt - 1)], Nt, nx, nw)  # :300-314
    g5 = _toeplitz_lower([Ap[k] @ full["b5"] for k in range(Nt - 1)], Nt, nx, 1)
    out["Gamma_5"] = g5 @ np.ones((Nt, 1))                                                  # :317-332
    Ct, Et, Gt = _blkdiag(full["C"], Nt), _blkdiag(full["E"], Nt), _blkdiag(full["G"], Nt)
    out["L_x"] = Ct @ out["Phi_x"] if nx else np.zeros((ny * Nt, 0))                        # :186
    out["L_v"] = Ct @ out["Gamma_v"] + _blkdiag(Dv, Nt)                                     # :187
    out["L_omega"] = Ct @ out["Gamma_omega"] + _blkdiag(full["D4"], Nt)                     # :188
    out["L_5"] = Ct @ out["Gamma_5"] + np.tile(full["d5"], (Nt, 1))                         # :189
    out["H_x"] = -(Et @ out["Phi_x"] + Gt @ out["L_x"]) if nx else np.zeros((nc * Nt, 0))   # :237
    out["H_v"] = Et @ out["Gamma_v"] + _blkdiag(Fv, Nt) + Gt @ out["L_v"]                   # :238
    out["H_omega"] = -(Et @ out["Gamma_omega"] + _blkdiag(full["F4"], Nt) + Gt @ out["L_omega"])  # :239
    out["H_5"] = np.tile(full["f5"], (Nt, 1)) - (Et @ out["Gamma_5"] + Gt @ out["L_5"])     # :240
    return out


def slice_N_p(mat, rows_per_step, N_p):
    """``*_N_p`` variants are row-prefix views -- mld_evolution_matrices.py:246-250."""
    return mat[:N_p * rows_per_step, :]


def output_bytes(dims, Nt):
    """Dense bytes of the 12 matrices: 8 (nx+ny+nc) Nt (2 + nv Nt + nomega Nt) minus the x-columns when nx=0."""
    tot = 0
    for r in (dims["nx"], dims["ny"], dims["nc"]):
        tot += r * Nt * (dims["nx"] + dims["nv"] * Nt + dims["nomega"] * Nt + 1)
    return 8 * tot
