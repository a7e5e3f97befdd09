"""Host pre-stage of the product (csrc/prestage.cu: CAMBParams_Set + InitVars + k grids + l sampling, batched over host
threads) against the CPU oracle's tables for the same parameters.  No GPU needed.  Tolerances: grids and integer
layouts identical; thermo tables 1e-10 relative to the column's scale (the product's RECFAST right-hand side hoists a few
common factors, so the adaptive steps can round differently)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cosmomc_galileon_b200 import make_params, prestage  # noqa: E402

GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243)
LCDM = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.00064, H0=70.0)
NONU = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.0, H0=70.0)


def oracle_tables(P, lmax, **extra):
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    o = Oracle(lmax=lmax, **P, **extra)
    o.stage("background", "initvars", "qgrid")
    return tables_from_oracle(o, dict(P, **extra), lmax=lmax)


def compare(t, ref, tol_tab=1e-10):
    d, r = t.d, ref.d
    for n in ("grhom grhog grhor grhob grhoc grhornomass adotrad tau0 taurst taurend tau_maxvis tight_tau matter_verydom_tau epsw "
              "max_etak_scalar tauminn dlntau dlnam").split():
        assert d[n] == pytest.approx(r[n], rel=1e-12, abs=0), n
    for n in "nt nk nq nthermo Nu_mass_eigenstates Num_Nu_massive nqmax lmax ngal use_galileon WantLateTime".split():
        assert int(d[n]) == int(r[n]), n
    assert int(d["kmaxfile"]) == int(max(min(int(d["lmax"]), 3000) * 2.5, 2.0 * int(d["lmax"]))) + 1  # camb/bessels.f90:58
    assert np.array_equal(np.asarray(d["ls"]), np.asarray(r["ls"]))
    # grids: bit-identical
    for n in ("k", "q", "dq", "tau", "dtau"):
        assert np.array_equal(d[n], r[n]), n
    np.testing.assert_array_equal(np.asarray(d["tau_regions"], dtype=float), np.asarray(r["tau_regions"], dtype=float))
    # thermo and visibility tables
    for n in "cs2 dcs2 dotmu ddotmu dddotmu vis dvis ddvis expmmu opac dopac".split():
        scale = np.max(np.abs(r[n]))
        assert np.max(np.abs(d[n] - r[n])) <= tol_tab * scale, (n, np.max(np.abs(d[n] - r[n])) / scale)
    for n in ("grhormass", "nu_masses", "nu_q", "nu_int_kernel", "nu_tau_nonrelativistic", "nu_tau_massive"):
        np.testing.assert_allclose(np.ravel(d[n]), np.ravel(r[n]), rtol=1e-12, err_msg=n)
    np.testing.assert_allclose(np.ravel(d["nu_tau_notmassless"]), np.ravel(r["nu_tau_notmassless"]), rtol=1e-12)
    if int(r["Nu_mass_eigenstates"]):
        for n in "r1 p1 dr1 dp1 ddr1 ddp1".split():
            np.testing.assert_allclose(d["nu_" + n], r["nu_" + n], rtol=1e-13, atol=1e-15, err_msg=n)
    if int(r["use_galileon"]):
        for n in "c2 c3 c4 c5 cG h0 orad".split():
            assert d[n] == pytest.approx(r[n], rel=1e-14), n
        np.testing.assert_array_equal(d["gal_a"], r["gal_a"])
        np.testing.assert_allclose(d["gal_h"], r["gal_h"], rtol=1e-12)
        np.testing.assert_allclose(d["gal_x"], r["gal_x"], rtol=1e-10)


@pytest.mark.parametrize("name,P,lmax,extra", [
    ("galileon", GAL, 2500, {}),
    ("lcdm", LCDM, 2500, {}),
    ("massless", NONU, 1200, {}),
    ("lensed_boost2", GAL, 5000, dict(DoLensing=1, AccuracyBoost=2.0)),
    ("three_eigenstates", dict(LCDM, omnuh2=0.0015), 1200, dict(Nu_mass_eigenstates=3, Num_Nu_massive=3, Num_Nu_massless=0.046)),
    ("galileon_three_eigenstates", GAL, 800, dict(Nu_mass_eigenstates=3, Num_Nu_massive=3, Num_Nu_massless=0.046)),
])
def test_prestage_tables_against_oracle(name, P, lmax, extra):
    pextra, oextra = dict(extra), dict(extra)
    if extra.get("Nu_mass_eigenstates", 1) == 3:  # a normal-hierarchy-like split (camb/inidriver.F90:143-165)
        fr = [0.2, 0.3, 0.5]
        pextra.update(Nu_mass_numbers=[1, 1, 1], Nu_mass_fractions=fr)
        for i in range(3):
            oextra["Nu_mass_numbers%d" % (i + 1)] = 1
            oextra["Nu_mass_fractions%d" % (i + 1)] = fr[i]
    ref = oracle_tables(P, lmax, **oextra)
    p = make_params(Max_l=lmax, Max_eta_k=2.0 * lmax, **P, **pextra)
    tabs, status = prestage([p], nthreads=1)
    assert status == [0]
    compare(tabs[0], ref)


def test_prestage_batch_is_threaded_and_rejects_bad_points():
    """A batch over several host threads returns the same tables as one point at a time; a point whose Galileon
    background is unstable gets CAMB's error_background_evolution (6) without disturbing the others."""
    good = make_params(Max_l=600, Max_eta_k=1200.0, **GAL)
    bad = make_params(Max_l=600, Max_eta_k=1200.0, **dict(GAL, c2=+5.0))
    other = make_params(Max_l=600, Max_eta_k=1200.0, **LCDM)
    tabs, status = prestage([good, bad, other, good], nthreads=4)
    assert status[0] == 0 and status[2] == 0 and status[3] == 0 and status[1] == 6
    assert tabs[1] is None
    single, _ = prestage([good], nthreads=1)
    for n in ("cs2", "dotmu", "vis", "k", "q", "gal_h"):
        assert np.array_equal(tabs[0].d[n], single[0].d[n]) and np.array_equal(tabs[3].d[n], single[0].d[n]), n


def test_transfer_grids_against_oracle():
    """CP%WantTransfer in the product pre-stage: the transfer times tautf (camb/cmbmain.f90:783-793), the wavenumbers
    InitTransfer appends beyond the source grid (:508-632) and the earlier thermo-table start (maxq, :757-762)
    are identical to the oracle's."""
    from oracle_py import Oracle
    from cosmomc_galileon_b200 import make_params, prestage
    GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
               c4=-1.03411, cG=0.001188243)
    Z = (2.0, 0.5, 0.0)
    for kmax, kpl in ((1.0, 0), (0.6, 5)):
        o = Oracle(lmax=800, **GAL)
        o.want_transfer(redshifts=Z, kmax=kmax, k_per_logint=kpl)
        o.stage("background", "initvars")
        To = o.transfer()
        p = make_params(Max_l=800, Max_eta_k=1600.0, WantTransfer=1, transfer_kmax=kmax, transfer_k_per_logint=kpl,
                        transfer_num_redshifts=len(Z), transfer_redshifts=Z, **GAL)
        tabs, st = prestage([p], nthreads=1)
        assert st == [0]
        t = tabs[0].d
        nk = o.geti("nk")
        assert t["nk"] == nk and t["WantTransfer"] == 1 and t["H0"] == GAL["H0"]
        assert t["nk_transfer"] == To["q_trans"].size - nk > 0
        assert np.max(np.abs(t["k_transfer"] / To["q_trans"][nk:] - 1)) < 1e-14
        assert np.max(np.abs(t["tautf"] / To["tautf"] - 1)) < 1e-12
        assert t["tauminn"] == o.get("tauminn")
    # out-of-order redshifts and the unsupported high-precision mode are rejected with error_unsupported_params
    bad = make_params(Max_l=800, Max_eta_k=1600.0, WantTransfer=1, transfer_num_redshifts=2, transfer_redshifts=(0.0, 1.0), **GAL)
    hp = make_params(Max_l=800, Max_eta_k=1600.0, WantTransfer=1, transfer_high_precision=1, **GAL)
    assert prestage([bad, hp], nthreads=1)[1] == [5, 5]
