"""CPU: pins the oracle (numpy restatement) against the golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden.py) and against the known-answer vector of SURVEY.md Appendix A."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden
from oracle import mld as omld, condense as oc, lsim as ol, ref_shim

EVO = ("Phi_x", "Gamma_v", "Gamma_omega", "Gamma_5", "L_x", "L_v", "L_omega", "L_5", "H_x", "H_v", "H_omega", "H_5")


def _mats(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


@pytest.mark.parametrize("case", golden_cases("condense"))
def test_condense_matches_reference_golden(case):
    g = load_golden("condense", case)
    full, d, vt = omld.complete(_mats(g), nu_l=int(g["nu_l"]))
    dims_ref = dict(zip(("nx", "nu", "ndelta", "nz", "nmu", "nomega", "ny", "nc"), g["dims"].tolist()))
    assert {k: d[k] for k in dims_ref} == dims_ref
    assert [t == "b" for t in vt] == g["var_type_v"].tolist()
    out = oc.condense(full, d, int(g["Nt"]))
    for name in EVO:
        ref = g["out_" + name]
        assert out[name].shape == ref.shape, name
        if ref.size:
            np.testing.assert_allclose(out[name], ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()), err_msg=name)
    assert sum(g["out_" + n].nbytes for n in EVO) == oc.output_bytes(d, int(g["Nt"]))


def test_appendix_a_known_answer():
    g = load_golden("condense", "appendixA")
    np.testing.assert_allclose(g["out_Phi_x"].ravel(), [1, 0.99, 0.9801, 0.970299], atol=1e-15)
    np.testing.assert_allclose(g["out_Gamma_5"].ravel(), [0, 0.1, 0.199, 0.29701], atol=1e-15)
    np.testing.assert_allclose(g["out_H_5"].ravel(), [65, -50, 64.9, -49.9, 64.801, -49.801, 64.70299, -49.70299],
                               atol=1e-12)
    Hv = g["out_H_v"]
    assert Hv.shape == (8, 12) and Hv[0, 1] == -1 and Hv[2, 0] == 0.5 and Hv[6, 6] == 0.5 and Hv[7, 11] == -1
    assert abs(Hv[6, 0] - 0.49005) < 1e-15


@pytest.mark.parametrize("case", golden_cases("lsim"))
def test_lsim_matches_reference_golden(case):
    g = load_golden("lsim", case)
    full, d, vt = omld.complete(_mats(g), nu_l=int(g["nu_l"]))
    for t in range(g["x"].shape[0]):
        x1, y, cons = ol.lsim_k(full, g["x"][t], g["u"][t], g["delta"][t], g["z"][t], g["mu"][t], g["w"][t])
        np.testing.assert_allclose(x1, g["x1"][t], rtol=0, atol=1e-12 * max(1.0, np.abs(g["x1"][t]).max(initial=0)))
        np.testing.assert_allclose(y, g["y"][t], rtol=0, atol=1e-12 * max(1.0, np.abs(g["y"][t]).max(initial=0)))
        assert np.array_equal(cons.astype(bool), g["cons"][t].astype(bool))


def test_dewh_closed_form_matches_survey_values():
    A, B1, B4, b5 = ol.dewh_scalars(ol.DEWH_PARAMS)
    assert abs(A - 0.997037112790) < 1e-12 and abs(B1 - 4.298192277481) < 1e-11
    assert abs(B4 + 179.733208275) < 1e-8 and abs(b5 - 0.074072180249) < 1e-12


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted (GPU box)")
def test_restatement_against_live_reference():
    rng = np.random.default_rng(5)
    mats = dict(A=rng.standard_normal((2, 2)) * 0.4, B1=rng.standard_normal((2, 1)), B4=rng.standard_normal((2, 1)),
                b5=rng.standard_normal((2, 1)), E=rng.standard_normal((3, 2)), F1=rng.standard_normal((3, 1)),
                Psi=-np.eye(3), f5=rng.standard_normal((3, 1)))
    ref, dims, _ = ref_shim.reference_condense(mats, 5, 6, bin_dims=dict(nu_l=1))
    full, d, vt = omld.complete(mats, nu_l=1)
    out = oc.condense(full, d, 6)
    for k in ref:
        np.testing.assert_allclose(out[k], ref[k], atol=1e-13)
