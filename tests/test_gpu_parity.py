"""GPU parity tests (run on the B200 box): every stage of the CUDA path through the C ABI against the oracle on
the same inputs, and against the committed golden fixtures.  Tolerances are the north-star ones: 1e-6 relative
(max-norm) on transfer functions / sources, 1e-5 relative on C_l."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_TRANSFER = 1e-6
TOL_CL = 1e-5
# per-element bound on the sources: relative for every element above 1e-6 of its source's maximum, absolute (in units of
# that floor) below; measured on the default (FMA-contracting) build: 8e-10 (lmax 800) ... 2.5e-6 (Galileon, lmax 2500,
# set by elements right at the floor: an absolute error of 2.5e-12 of the maximum)
TOL_ELEMENT = 1e-5
GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243)
LCDM = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.00064, H0=70.0)


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def colmax(S, So):
    """Sources [t][s][k]: worst max-norm relative error over the (source, k) columns -- every k mode is its own
    integration with its own amplitude (high-k columns are orders of magnitude below the low-k ones), so this is far
    stricter than one max-norm over the whole array."""
    num = np.max(np.abs(S - So), axis=0)
    den = np.max(np.abs(So), axis=0)
    ok = den > 0
    return float(np.max(num[ok] / den[ok])) if np.any(ok) else 0.0


def elemmax(S, So, floor=1e-6):
    """Per-element error: relative where |So| > floor * max|So| of its source, absolute (in units of that floor)
    elsewhere."""
    worst = 0.0
    for s in range(So.shape[1]):
        m = np.max(np.abs(So[:, s]))
        if m == 0:
            continue
        worst = max(worst, float(np.max(np.abs(S[:, s] - So[:, s]) / np.maximum(np.abs(So[:, s]), floor * m))))
    return worst


@pytest.fixture(scope="module")
def gpu():
    from cosmomc_galileon_b200 import GalCamb
    g = GalCamb(0)
    tmpl = np.fromfile(os.path.join(GOLD, "highL_template.f64")).reshape(7999, 4).T.copy()  # TT, EE, TE, PP
    g.set_template(tmpl)
    yield g


def run_gpu(g, t):
    g.InitSpherBessels(t.ls, t.d["kmaxfile"], t.d["AccuracyBoost"])
    g.set_models([t])
    g.CalcScalarSources()
    g.SourceToTransfers()
    g.ClTransferToCl()


@pytest.mark.parametrize("P", [LCDM, GAL], ids=["lcdm", "galileon"])
def test_full_path_against_oracle(gpu, P):
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    o = Oracle(**P)
    o.run()
    t = tables_from_oracle(o, P)
    run_gpu(gpu, t)
    S, So = gpu.Src(), o.Src()
    for s in range(S.shape[1]):
        assert relmax(S[:, s], So[:, s]) < TOL_TRANSFER
    # per (source, k) column, and per element (relative down to 1e-6 of the source's maximum, absolute below)
    assert colmax(S, So) < TOL_TRANSFER
    assert elemmax(S, So) < TOL_ELEMENT, elemmax(S, So)
    assert relmax(gpu.Delta_p_l_k(), o.Delta()) < TOL_TRANSFER
    icl, cl = gpu.Cls()
    assert np.max(np.abs(cl / o.Cl() - 1)) < TOL_CL
    assert np.max(np.abs(icl / o.iCl() - 1)) < TOL_CL
    # same adaptive-step history as the CPU restatement: the number of RHS evaluations agrees (exactly when the
    # library is built with -fmad=false; FMA contraction may flip a borderline accept/reject, hence the 1e-4 slack)
    assert abs(gpu.counters()["n_derivs"] / o.get("n_derivs") - 1) < 1e-4
    assert gpu.counters()["n_iter"] == o.get("n_iter")


@pytest.mark.parametrize("kw,lmax", [(dict(DoLensing=1, AccuracyBoost=2), 3000), (dict(DoLensing=1, DoLateRadTruncation=0), 2650)],
                         ids=["lensing_accuracyboost2", "cosmomc_settings"])
def test_three_source_variants_against_oracle(gpu, kw, lmax):
    """BASELINE configs[2] / configs[4] flavours: lensing source (S=3, six C_l types, Limber substitution in the LOS
    stage), AccuracyBoost=2 (nqmax=4, halved steps, tol1=1e-4/e) and the CosmoMC setting DoLateRadTruncation=F
    (SURVEY.md section 3.2).  lmax reduced to keep the single-model run short."""
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    P = dict(GAL)
    P.update(kw)
    o = Oracle(lmax=lmax, **P)
    o.run()
    t = tables_from_oracle(o, P, lmax=lmax)
    run_gpu(gpu, t)
    S, So = gpu.Src(), o.Src()
    assert S.shape[1] == 3
    for s in range(3):
        assert relmax(S[:, s], So[:, s]) < TOL_TRANSFER
    assert relmax(gpu.Delta_p_l_k(), o.Delta()) < TOL_TRANSFER
    icl, cl = gpu.Cls()
    assert icl.shape[0] == 6
    assert np.max(np.abs(icl[:3] / o.iCl()[:3] - 1)) < TOL_CL
    assert np.max(np.abs(cl[:4] / o.Cl()[:4] - 1)) < TOL_CL  # TT, EE, TE, phi-phi
    for t_ in (4, 5):  # phi-T and phi-E cross spectra change sign: compare on the max-norm
        assert relmax(cl[t_], o.Cl()[t_]) < TOL_CL
    # stage 5: lensed spectra from the device-resident C_l (camb/lensing.f90:107-528)
    o.stage("lensing")
    gpu.lens_Cls()
    check_lensed(gpu.Cl_lensed(), o.Cl_lensed(), TOL_CL)
    assert gpu.Cl_lensed().shape[1] == o.geti("lmax_lensed") - 1
    # and through the one-call batch entry CosmoMC would use with DoLensing = T
    tau = t.d["tau"].copy()
    tau[40] = tau[39]  # a point dverk rejects (see test_failed_point_does_not_sink_the_batch)
    cl_b, lensed_b = gpu.batch_lensed_cls([t, t.copy(tau=tau)])
    check_lensed(lensed_b[0], o.Cl_lensed(), TOL_CL)
    assert np.max(np.abs(cl_b[0, :2] / o.Cl()[:2] - 1)) < TOL_CL
    assert gpu.status(0) == 0 and gpu.status(1) == 4 and np.all(np.isnan(lensed_b[1]))
    # fast-parameter update (CAMB_TransfersToPowers, camb/camb.f90:87): new A_s, n_s on the resident transfer functions
    o.set(ScalarPowerAmp=2.3e-9, an=0.95)
    o.stage("cls", "lensing")
    gpu.set_primordial(0, 2.3e-9, 0.95)
    n_before = gpu.evolve_stats()["rhs_executed"]
    gpu.TransfersToPowers(do_lensing=True)
    assert gpu.evolve_stats()["rhs_executed"] == n_before  # no ODE work
    _, cl2 = gpu.Cls(0)
    assert np.max(np.abs(cl2[:2] / o.Cl()[:2] - 1)) < TOL_CL
    check_lensed(gpu.Cl_lensed(0), o.Cl_lensed(), TOL_CL)
    # the lensing amplitudes CosmoMC sets before every call (source/Calculator_CAMB.f90:130-131): ALens scales C_phiphi and,
    # by its root, the phi-T / phi-E spectra (camb/cmbmain.f90:2254-2259); ALens_Fiducial puts the scaled template potential
    # in place of the computed one (camb/lensing.f90:252-257)
    for al, alf in ((1.3, 0.0), (1.0, 1.1)):
        o.set(ALens=al, ALens_Fiducial=alf)
        o.stage("cls", "lensing")
        gpu.set_alens(0, al, alf)
        gpu.TransfersToPowers(do_lensing=True)
        _, cl3 = gpu.Cls(0)
        assert np.max(np.abs(cl3[3] / o.Cl()[3] - 1)) < TOL_CL
        assert np.max(np.abs(cl3[3] / cl2[3] / al - 1)) < 1e-12 and relmax(cl3[4], np.sqrt(al) * cl2[4]) < 1e-12
        check_lensed(gpu.Cl_lensed(0), o.Cl_lensed(), TOL_CL)
    assert gpu.evolve_stats()["rhs_executed"] == n_before
    o.set(ALens=1.0, ALens_Fiducial=0.0)
    gpu.set_alens(0, 1.0, 0.0)


def test_config3_lensed_accuracyboost2_lmax5000(gpu):
    """BASELINE configs[2] at full size: Galileon lensed C_l, AccuracyBoost = 2, lmax = 5000 (557 source k, 1238 output
    times, 10814 integration k, 135 multipoles, 3 sources, every k integrated to tau0), all five stages."""
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    P = dict(GAL, DoLensing=1, AccuracyBoost=2)
    o = Oracle(lmax=5000, **P)
    o.run()
    o.stage("lensing")
    t = tables_from_oracle(o, P, lmax=5000)
    run_gpu(gpu, t)
    gpu.lens_Cls()
    S, So = gpu.Src(), o.Src()
    for s in range(3):
        assert relmax(S[:, s], So[:, s]) < TOL_TRANSFER
    assert relmax(gpu.Delta_p_l_k(), o.Delta()) < TOL_TRANSFER
    icl, cl = gpu.Cls()
    assert np.max(np.abs(cl[:4] / o.Cl()[:4] - 1)) < TOL_CL
    check_lensed(gpu.Cl_lensed(), o.Cl_lensed(), TOL_CL)
    assert gpu.Cl_lensed().shape[1] == 4899
    print("config 3 stage ms:", gpu.timings())


def check_lensed(L, Lo, tol):
    assert L.shape == Lo.shape
    for t_ in (0, 1, 2):  # TT, EE, BB are positive
        assert np.max(np.abs(L[t_] / Lo[t_] - 1)) < tol, t_
    # TE changes sign: max-norm over a sliding window of 50 multipoles (the spectrum falls by 1e4 over the range)
    n = L.shape[1] // 50 * 50
    a, b = L[3, :n].reshape(-1, 50), Lo[3, :n].reshape(-1, 50)
    assert np.max(np.max(np.abs(a - b), axis=1) / np.max(np.abs(b), axis=1)) < tol


def test_lensing_stage_alone_high_l(gpu):
    """BASELINE configs[2] geometry for the lensing stage (lmax 5000, AccuracyBoost 2: 1625 angles in 13 CTAs per
    model, the Max_l > 3500 step refinement, template extrapolation to l = 5650) without paying for the AccuracyBoost=2
    ODE: both sides get the C_l of an AccuracyBoost=1 oracle run.  Then the properties the transform has at any size:
    linearity in the unlensed CMB spectra at fixed potential, and a realistic-normalisation failure stays in its slot."""
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    P = dict(GAL)
    P.update(DoLensing=1)
    oA = Oracle(lmax=5000, **P)
    oA.run()
    Cl = oA.Cl()
    P2 = dict(P)
    P2.update(AccuracyBoost=2)
    oB = Oracle(lmax=5000, **P2)
    oB.stage("background", "initvars", "qgrid")
    oB.put("Cl", Cl)
    oB.stage("lensing")
    Lo = oB.Cl_lensed()
    assert oB.geti("lmax_lensed") == 4900
    t = tables_from_oracle(oB, P2, lmax=5000)
    gpu.InitSpherBessels(t.ls, t.d["kmaxfile"], t.d["AccuracyBoost"])
    Cl2 = Cl.copy()
    Cl2[:3] *= 0.6
    Cl3 = Cl.copy()
    Cl3[3] *= 1e3  # lensing potential far too large: error_norm_lensing (camb/lensing.f90:230-235)
    Cl0 = Cl.copy()
    Cl0[3:] = 0  # no lensing potential: the transform must return the unlensed spectra and no B modes
    gpu.set_models([t, t.copy(), t.copy(), t.copy()])
    for i, c in enumerate((Cl, Cl2, Cl3, Cl0)):
        gpu.put_Cls(c, i)
    gpu.lens_Cls()
    L0 = gpu.Cl_lensed(3)
    nl = L0.shape[1]
    assert np.max(np.abs(L0[0] / Cl[0, :nl] - 1)) < 1e-11 and np.max(np.abs(L0[1] / Cl[1, :nl] - 1)) < 1e-11
    assert np.max(np.abs(L0[2])) < 1e-24 and np.max(np.abs(L0[3] - Cl[2, :nl])) < 1e-11 * np.max(np.abs(Cl[2]))
    L = gpu.Cl_lensed(0)
    check_lensed(L, Lo, TOL_CL)
    np.testing.assert_allclose(gpu.Cl_lensed(1), 0.6 * L, rtol=1e-11)
    assert gpu.status(0) == 0 and gpu.status(1) == 0 and gpu.status(2) == 8
    assert np.all(np.isnan(gpu.Cl_lensed(2)))


def test_bessel_tables_against_oracle(gpu):
    from oracle_py import Oracle
    o = Oracle(**LCDM)
    o.stage("background", "initvars", "bessels")
    ls = o.arr("ls").astype(np.int32)
    gpu.InitSpherBessels(ls, o.geti("kmaxfile"), 1.0)
    xs, ajl, ajlpr = gpu.bessel_tables()
    assert np.array_equal(xs, o.arr("bess_x"))  # the x grid is bit-identical (camb/bessels.f90:68-76)
    assert relmax(ajl.ravel(), o.arr("ajl")) < 1e-11
    assert relmax(ajlpr.ravel(), o.arr("ajlpr")) < 1e-7  # second derivatives amplify the 1-ulp sin/cos differences


def test_los_stage_alone_on_oracle_sources(gpu):
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    o = Oracle(**GAL)
    o.run()
    t = tables_from_oracle(o, GAL)
    gpu.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    gpu.set_models([t])
    gpu.put_Src(o.Src())
    gpu.SourceToTransfers()
    assert relmax(gpu.Delta_p_l_k(), o.Delta()) < 1e-10  # same operands: only the summation order differs
    # linearity of the LOS stage: scaling the sources scales the transfers
    gpu.put_Src(3.0 * o.Src())
    gpu.SourceToTransfers()
    assert relmax(gpu.Delta_p_l_k(), 3.0 * o.Delta()) < 1e-10


def test_golden_fixture_and_batch(gpu):
    """config-2 table fixture -> C_l equal to the committed golden; a batch of replicas gives identical results
    in every slot (no cross-talk between work items), and the one-call ABI agrees with the staged calls."""
    from cosmomc_galileon_b200 import ModelTables
    t = ModelTables.load(os.path.join(GOLD, "config2_tables.npz"))
    gold = np.load(os.path.join(GOLD, "config2_golden.npz"))
    gpu.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    B = 40  # not a multiple of 32: ragged last warp per k index
    out = gpu.batch_cls([t.copy() for _ in range(B)])
    assert np.max(np.abs(out[0] / gold["Cl"] - 1)) < TOL_CL
    # no cross-talk: replicas that share a kind of warp are bit-identical.  Slots 0..31 fill one warp per k index, slots
    # 32..39 sit in a warp of 8 items whose stage combinations are formed by combine_group() (idle lanes help): same
    # sums in the same order, but a different instruction schedule (FMA contraction) -> agreement at round-off level
    for b in range(1, 32):
        assert np.array_equal(out[b], out[0])
    for b in range(33, B):
        assert np.array_equal(out[b], out[32])
    assert np.max(np.abs(out[32] / out[0] - 1)) < 1e-8
    S = gpu.Src(B - 1)[::16]
    assert relmax(S, gold["Src_t"]) < TOL_TRANSFER
    D = gpu.Delta_p_l_k(7)[::32]
    assert relmax(D, gold["Delta_q"]) < TOL_TRANSFER


def test_heterogeneous_batch(gpu):
    """A batch of DIFFERENT cosmologies (LCDM, the Galileon point, a shifted Galileon point: different k and time
    grids, mixed use_galileon inside one warp) -- every slot must match its own oracle run."""
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    GAL2 = dict(GAL, H0=78.2, ombh2=0.02225, optical_depth=0.08, an=0.955)  # (the test point sits close to the no-ghost boundary: larger shifts are rejected by galileon.cc:209-242)
    runs = []
    for P in (LCDM, GAL, GAL2):
        o = Oracle(**P)
        o.run()
        runs.append((o, tables_from_oracle(o, P)))
    assert len({t.d["nk"] for _, t in runs} | {t.d["nt"] for _, t in runs}) > 2  # the grids really differ
    t0 = runs[0][1]
    gpu.InitSpherBessels(t0.ls, max(t.d["kmaxfile"] for _, t in runs), 1.0)
    out = gpu.batch_cls([t for _, t in runs])
    for i, (o, t) in enumerate(runs):
        assert gpu.status(i) == 0
        assert np.max(np.abs(out[i] / o.Cl()[:3] - 1)) < TOL_CL
        assert relmax(gpu.Src(i), o.Src()) < TOL_TRANSFER
        assert relmax(gpu.Delta_p_l_k(i), o.Delta()) < TOL_TRANSFER


def test_distinct_cosmology_batch(gpu):
    """BASELINE configs[3] in small: 48 DIFFERENT parameter points drawn like SURVEY.md 8d's config 4 (seeded Gaussians
    around the config-2 point, stability rejects redrawn).  Their k grids, switch times and adaptive step sequences
    differ, so the lanes of a warp fall out of step: this is the test of the stage alignment, the shared output
    derivative and the deferred output evaluation of K1.  Every 6th point is checked against its own oracle run."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    from gpu_hetero_probe import draw, prestage
    from oracle_py import Oracle
    rng = np.random.default_rng(7)
    tabs, pars = [], []
    while len(tabs) < 48:
        P = draw(rng, 1.0)
        t = prestage(P)
        if t is not None:
            tabs.append(t)
            pars.append(P)
    gpu.InitSpherBessels(tabs[0].ls, 6251, 1.0)
    out = gpu.batch_cls(tabs)
    assert all(gpu.status(i) == 0 for i in range(48))
    st = gpu.evolve_stats()
    assert st["outputs_from_k1"] > 0.9 * gpu.counters()["n_out"]  # nearly every output shares the next step's k1
    n_alg = 0.0
    for i in range(0, 48, 6):
        o = Oracle(**pars[i])
        o.run()
        assert np.max(np.abs(out[i] / o.Cl()[:3] - 1)[:2]) < TOL_CL
        assert relmax(out[i][2], o.Cl()[2]) < TOL_CL
        assert relmax(gpu.Src(i), o.Src()) < TOL_TRANSFER
        n_alg += o.get("n_derivs")
    assert n_alg > 0


def test_distinct_cosmology_batch_in_batch_mode():
    """The same kind of batch through the lane-per-item form of K1 (csrc/evolve_batch.cu) in its lock-step launch shape
    (4-warp CTAs), which the launcher only picks for large batches: forced here with GALCAMB_OCTET_ITEMS=0 (lane-per-item
    form whatever the size) and GALCAMB_LATENCY_ITEMS=0 (lock-step CTAs) in a fresh process (the knobs are read once per
    process).  40 distinct points = one full warp and one warp of 8 items (cooperative combine) per k index."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GALCAMB_LATENCY_ITEMS="0", GALCAMB_OCTET_ITEMS="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_hetero_probe.py"), "40"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(re.findall(r"failed slots 0;", r.stdout)) == 2, r.stdout
    m = re.search(r"vs the oracle: ([0-9.e+-]+)", r.stdout)
    assert m and float(m.group(1)) < TOL_CL, r.stdout


def test_both_forms_of_k1_on_the_golden_fixture():
    """The committed config-2 fixture through each form of K1 (eight lanes per item / one lane per item, one-warp CTAs),
    forced by GALCAMB_OCTET_ITEMS in fresh processes: both reproduce the golden sources and C_l."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for knob in ("1000000000", "0"):
        env = dict(os.environ, GALCAMB_OCTET_ITEMS=knob)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_k1_probe.py"), "3"], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        m = re.search(r"Cl err ([0-9.e+-]+) Src err ([0-9.e+-]+) identical=True", r.stdout)
        assert m, r.stdout
        assert float(m.group(1)) < TOL_CL and float(m.group(2)) < TOL_TRANSFER, r.stdout


NU3 = dict(Nu_mass_eigenstates=3, Num_Nu_massive=3, Num_Nu_massless=0.046)
FR3 = [0.2, 0.3, 0.5]


def oracle_three_eigenstates(P, lmax, **kw):
    from oracle_py import Oracle
    o = Oracle(lmax=lmax, **P, **NU3, **kw)
    for i in range(3):
        o.set(**{"Nu_mass_numbers%d" % (i + 1): 1, "Nu_mass_fractions%d" % (i + 1): FR3[i]})
    return o


@pytest.mark.parametrize("P,lmax", [(dict(LCDM, omnuh2=0.0015), 1200), (GAL, 1000)], ids=["lcdm", "galileon"])
def test_three_mass_eigenstates_against_oracle(gpu, P, lmax):
    """Three massive eigenstates with different masses (camb/inidriver.F90:143-165, the hierarchies CosmoMC sets through
    CAMB_SetNeutrinoHierarchy, source/Calculator_CAMB.f90:139): three momentum-node blocks in the state vector, three
    sets of switch times, the per-eigenstate fluid approximation."""
    from tables_from_oracle import tables_from_oracle
    o = oracle_three_eigenstates(P, lmax)
    o.run()
    assert o.geti("Nu_mass_eigenstates") == 3
    t = tables_from_oracle(o, dict(P, **NU3), lmax=lmax)
    assert len(t.d["nu_masses"]) == 3 and len(set(np.round(t.d["nu_masses"], 6))) == 3
    run_gpu(gpu, t)
    S, So = gpu.Src(), o.Src()
    assert colmax(S, So) < TOL_TRANSFER
    assert relmax(gpu.Delta_p_l_k(), o.Delta()) < TOL_TRANSFER
    _, cl = gpu.Cls()
    assert np.max(np.abs(cl / o.Cl() - 1)) < TOL_CL
    assert abs(gpu.counters()["n_derivs"] / o.get("n_derivs") - 1) < 1e-4


def test_three_eigenstates_through_the_batch_kernel():
    """The same through the lane-per-item form of K1 (forced with GALCAMB_OCTET_ITEMS=0 in a fresh process)."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GALCAMB_OCTET_ITEMS="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_nu3_probe.py"), "800"], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"colmax ([0-9.e+-]+) cl ([0-9.e+-]+) n_derivs gpu ([0-9]+) oracle ([0-9]+)", r.stdout)
    assert m and float(m.group(1)) < TOL_TRANSFER and float(m.group(2)) < TOL_CL, r.stdout
    assert abs(float(m.group(3)) / float(m.group(4)) - 1) < 1e-4


def test_parameters_to_cls_without_the_oracle_in_the_loop(gpu):
    """galcamb_params_to_cls_: parameters in, C_l out, every table built by the product's own pre-stage
    (csrc/prestage.cu).  Compared with the oracle's C_l for the same parameters; the unstable point comes back as NaN."""
    from oracle_py import Oracle
    from cosmomc_galileon_b200 import make_params
    lmax = 1500
    pts = [GAL, LCDM, dict(GAL, c2=+5.0), dict(GAL, ombh2=0.0225, omch2=0.1160)]
    params = [make_params(Max_l=lmax, Max_eta_k=2.0 * lmax, **P) for P in pts]
    out = gpu.params_to_cls(params)
    assert np.all(np.isnan(out[2]))
    for i in (0, 1, 3):
        o = Oracle(lmax=lmax, **pts[i])
        o.run()
        assert np.max(np.abs(out[i] / o.Cl() - 1)) < TOL_CL, i


def test_device_background_against_host(gpu):
    """galcamb_set_prestage_device_: the Galileon backgrounds of a batch integrated on the device (one lane per cosmology,
    csrc/galileon_host.cu) against the same integrator on the host threads -- same tables (1e-12; the last bit of
    log / exp / pow is all that differs), same rejections, same C_l."""
    from cosmomc_galileon_b200 import make_params, prestage
    lmax = 800
    pts = [GAL, dict(GAL, c2=+5.0), dict(GAL, ombh2=0.0225, omch2=0.1160), LCDM, dict(GAL, H0=76.5, c3=-3.2),
           dict(GAL, omnuh2=0.0)]
    params = [make_params(Max_l=lmax, Max_eta_k=2.0 * lmax, **P) for P in pts]
    th, sh = prestage(params, nthreads=2, copy=True)
    cl_h = gpu.params_to_cls(params)
    gpu.set_prestage_device(0)
    try:
        td, sd = prestage(params, nthreads=2, copy=True)
        cl_d = gpu.params_to_cls(params)
    finally:
        gpu.set_prestage_device(-1)
    assert sh == sd and sh[1] != 0 and sh[0] == 0
    for a, b in zip(th, td):
        if a is None:
            assert b is None
            continue
        # (cs2 is left out: the reference's baryon-temperature recursion of inithermo amplifies a 2e-13 change of H0 to
        # O(1) differences in some 2600 rows of that table -- rounding noise that two runs only share bit for bit; the
        # C_l below are what it feeds)
        for key in ("gal_a", "gal_h", "gal_x", "tau", "vis", "dotmu", "k", "q"):
            x, y = np.asarray(a.d[key]), np.asarray(b.d[key])
            assert x.shape == y.shape, key
            if x.size:
                assert np.max(np.abs(x - y)) <= 2e-11 * np.max(np.abs(x)), key
        for key in ("c2", "c3", "c4", "c5", "cG", "tau0"):
            assert abs(a.d[key] - b.d[key]) <= 1e-13 * max(1.0, abs(a.d[key])), key
    ok = ~np.isnan(cl_h[:, 0, 0])
    assert np.array_equal(ok, ~np.isnan(cl_d[:, 0, 0]))
    assert np.max(np.abs(cl_d[ok] / cl_h[ok] - 1)) < 1e-8


def test_several_gpus_from_one_process(gpu):
    """The two multi-device entry points of the C ABI (parameter-batch sharding; one spectrum with its k-modes dealt over
    the devices and NCCL all-reduces of Src and iCl) on every device count the box offers, against the single-device
    result: same C_l to rounding (the sums over q are split, so not bit for bit)."""
    from cosmomc_galileon_b200 import make_params
    lmax = 800
    pts = [GAL, dict(GAL, ombh2=0.0225, omch2=0.1160), LCDM, dict(GAL, c2=+5.0), dict(GAL, H0=77.5)]
    params = [make_params(Max_l=lmax, Max_eta_k=2.0 * lmax, **P) for P in pts]
    ref = gpu.params_to_cls(params)
    one, ms1 = gpu.multi_single_cls(1, params[0])
    np.testing.assert_allclose(one, ref[0], rtol=1e-12)
    ndev = gpu.device_count()
    for G in [n for n in (1, 2, 4, 8) if n <= ndev]:
        out = gpu.multi_params_to_cls(G, params)
        assert np.all(np.isnan(out[3]))
        for i in (0, 1, 2, 4):
            np.testing.assert_allclose(out[i], ref[i], rtol=1e-12)
        single, ms = gpu.multi_single_cls(G, params[0])
        np.testing.assert_allclose(single, ref[0], rtol=1e-9)
        assert ms["call"] > 0 and ms["evolve"] > 0
    # the module fixture's context is still the calling thread's one
    assert gpu.status(0) in (0, 4)


def test_exact_step_history_without_fma_contraction():
    """Built with -fmad=false (GALCAMB_FMAD=false, own object directory) both forms of K1 take exactly the accept/reject
    decisions of the CPU oracle (built -ffp-contract=off): the number of right-hand-side evaluations of the reference
    algorithm is EQUAL, not just close.  The variant library is built on the box if nvcc is there; skipped otherwise."""
    import re
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "cosmomc_galileon_b200", "libgalcamb_b200_nofma.so")
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):  # incremental: a no-op when the variant is fresh
        env = dict(os.environ, GALCAMB_FMAD="false", GALCAMB_VARIANT="nofma")
        subprocess.check_call([sys.executable, os.path.join(root, "cosmomc_galileon_b200", "build.py")], env=env)
    elif not os.path.exists(lib):
        pytest.skip("no nvcc to build the -fmad=false variant")
    gold = np.load(os.path.join(GOLD, "config2_tables.npz"))
    meta = json.loads(bytes(gold["meta_json"]).decode()) if "meta_json" in gold.files else {}
    n_oracle = float(meta.get("n_derivs", 6360445.0))
    for knob in ("1000000000", "0"):
        env = dict(os.environ, GALCAMB_LIB=lib, GALCAMB_OCTET_ITEMS=knob)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_k1_probe.py"), "2"], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        m = re.search(r"'n_derivs': np.float64\(([0-9.]+)\)", r.stdout)
        assert m, r.stdout
        assert float(m.group(1)) == 2 * n_oracle, (knob, r.stdout)


# Matter transfer functions.  Rows of TransferData that are reproducible: the matter rows (cdm, b, tot,
# nonu, baryon-cdm velocity) and k/h; the photon / massless-neutrino / total-with-dark-energy / Weyl rows of the
# wavenumbers BEYOND the source grid carry the Galileon field perturbation, which at k ~ 1/Mpc is integrated in long free
# steps and is sensitive to last-bit differences (two CPU builds of the oracle, -ffp-contract=off vs -march=native,
# differ there by 2e-3 ... 8e-3 of the element, 1e-5 of the row's maximum over k; LCDM has no such rows).
MATTER_ROWS = [0, 1, 2, 6, 7, 12]


def _transfer_errors(T, To):
    a, b = T["TransferData"], To["TransferData"]
    assert a.shape == b.shape
    elem = float(np.max(np.abs(a[:, :, MATTER_ROWS] / b[:, :, MATTER_ROWS] - 1)))
    den = np.max(np.abs(b), axis=1, keepdims=True)
    den[den == 0] = 1.0
    return elem, float(np.max(np.abs(a - b) / den))


@pytest.mark.parametrize("P", [LCDM, GAL], ids=["lcdm", "galileon"])
def test_matter_transfer_against_oracle(gpu, P):
    """CP%WantTransfer through the product: transfer times and extra wavenumbers from the pre-stage, outtransf taps inside
    the time-step loop of K1 (camb/cmbmain.f90:1017-1068), the transfer-only wavenumbers (TransferOut, :1151-1201) and
    sigma_8 (camb/modules.f90:2304-2370), against the oracle; the C_l of the same call too."""
    from oracle_py import Oracle
    from cosmomc_galileon_b200 import make_params, prestage
    Z, kmax, lmax = (2.0, 0.5, 0.0), 1.0, 800
    o = Oracle(lmax=lmax, **P)
    o.want_transfer(redshifts=Z, kmax=kmax)
    o.run()
    To = o.transfer()
    p = make_params(Max_l=lmax, Max_eta_k=2.0 * lmax, WantTransfer=1, transfer_kmax=kmax, transfer_num_redshifts=len(Z),
                    transfer_redshifts=Z, **P)
    tabs, st = prestage([p])
    assert st == [0]
    run_gpu(gpu, tabs[0])
    T = gpu.transfer(0)
    assert T["q_trans"].size > o.geti("nk") and np.array_equal(T["q_trans"], To["q_trans"])
    elem, rowmax = _transfer_errors(T, To)
    assert elem < TOL_TRANSFER, elem                                 # per element, matter rows
    assert rowmax < (TOL_TRANSFER if not P.get("use_galileon") else 2e-5), rowmax  # every row, relative to its maximum over k
    assert np.max(np.abs(T["sigma8"] / To["sigma8"] - 1)) < 1e-9
    assert colmax(gpu.Src(), o.Src()) < TOL_TRANSFER
    assert np.max(np.abs(gpu.Cls()[1] / o.Cl() - 1)) < TOL_CL
    assert abs(gpu.counters()["n_derivs"] / o.get("n_derivs") - 1) < 1e-4


def test_matter_transfer_through_the_batch_kernel():
    """The same through the lane-per-item form of K1 (forced with GALCAMB_OCTET_ITEMS=0 in a fresh process)."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GALCAMB_OCTET_ITEMS="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_transfer_probe.py"), "3"], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"transfer err (\S+) sigma8 err (\S+) q_trans err (\S+) Src err (\S+) Cl err (\S+)", r.stdout)
    assert m, r.stdout
    te, s8, qe, se, ce = (float(x) for x in m.groups())
    assert te < 2e-5 and s8 < 1e-9 and qe == 0 and se < TOL_TRANSFER and ce < TOL_CL, r.stdout


def test_massless_neutrino_model(gpu):
    """No massive eigenstate (omnuh2 = 0 -> Num_Nu_massive = 0, camb/modules.f90:307-316): the neutrino blocks of the
    state vector and the table lookups disappear."""
    from oracle_py import Oracle
    from tables_from_oracle import tables_from_oracle
    P = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.0, H0=70.0)
    o = Oracle(lmax=1200, **P)
    o.run()
    assert o.geti("Num_Nu_massive") == 0
    t = tables_from_oracle(o, P, lmax=1200)
    run_gpu(gpu, t)
    assert relmax(gpu.Src(), o.Src()) < TOL_TRANSFER
    _, cl = gpu.Cls()
    assert np.max(np.abs(cl / o.Cl() - 1)) < TOL_CL


def test_failed_point_does_not_sink_the_batch(gpu):
    """Error behaviour: a point whose integration cannot proceed (here: a repeated output time, which dverk rejects
    like the reference's 'x = xend' error, camb/subroutines.f90:700-715) gets CAMB's error_evolution id and NaN C_l;
    the other points of the batch are unaffected."""
    from cosmomc_galileon_b200 import ModelTables
    t = ModelTables.load(os.path.join(GOLD, "config2_tables.npz"))
    gold = np.load(os.path.join(GOLD, "config2_golden.npz"))
    gpu.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    tau = t.d["tau"].copy()
    tau[40] = tau[39]
    bad = t.copy(tau=tau)
    out = gpu.batch_cls([t.copy(), bad, t.copy()])
    assert gpu.status(0) == 0 and gpu.status(2) == 0 and gpu.status(1) == 4
    assert np.all(np.isnan(out[1]))
    for i in (0, 2):
        assert np.max(np.abs(out[i] / gold["Cl"] - 1)) < TOL_CL


def test_primordial_rescaling_property(gpu):
    """C_l is linear in A_s (size-independent property of the K4 stage): doubling ScalarPowerAmp doubles iCl."""
    from cosmomc_galileon_b200 import ModelTables
    t = ModelTables.load(os.path.join(GOLD, "config2_tables.npz"))
    gpu.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    t2 = t.copy(ScalarPowerAmp=2 * t.d["ScalarPowerAmp"])
    gpu.set_models([t, t2])
    gpu.CalcScalarSources()
    gpu.SourceToTransfers()
    gpu.ClTransferToCl()
    i1, _ = gpu.Cls(0)
    i2, _ = gpu.Cls(1)
    np.testing.assert_allclose(i2, 2 * i1, rtol=1e-13)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
