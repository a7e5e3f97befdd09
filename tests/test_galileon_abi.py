"""The reference's own C ABI for the Galileon background and perturbation terms (include/galileon_abi.h; replacement of
camb/galileon.cc, host code of libgalcamb_b200.so) against the oracle's restatement of the same functions: background
integration with its own RKF45 + stability checks, spline accessors, every closed form, error statuses.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle_py import Oracle

GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243)
D = C.c_double
PD = C.POINTER(D)


def ref(x):
    return C.byref(D(x))


@pytest.fixture(scope="module")
def lib(built_cuda_lib):
    L = C.CDLL(built_cuda_lib)
    for f in ("GetX_", "GetH_", "ct_", "grhogal_", "gpresgal_", "Chigal_", "qgal_", "Pigal_", "dphisecond_", "pigalprime_"):
        getattr(L, f).restype = D
    return L


@pytest.fixture(scope="module")
def oracle():
    o = Oracle(template=False, **GAL)
    o.stage("background")
    return o


def call_arrays(L, o, c2=None, c3=None, c4=None, cG=None):
    g = o.get
    tabs = [np.ascontiguousarray(o.arr("nu_" + n)) for n in "r1 p1 dr1 dp1 ddr1 ddp1".split()]
    rc = L.galileon_set_nu_tables_(*[t.ctypes.data_as(PD) for t in tabs], C.byref(C.c_int(len(tabs[0]))), ref(g("nu_dlnam")))
    assert rc == 0
    grhormass = (D * 5)(g("grhormass1"))
    nu_masses = (D * 5)(g("nu_masses1"))
    args = [g("gal_orad"), g("gal_om"), g("gal_h0"), g("gal_c2") if c2 is None else c2, g("gal_c3") if c3 is None else c3,
            g("gal_c4") if c4 is None else c4, g("gal_cG") if cG is None else cG]
    return L.arrays_(*[ref(v) for v in args], grhormass, nu_masses, C.byref(C.c_int(1))), grhormass, nu_masses


def test_background_table_and_accessors(lib, oracle):
    rc, gm, nm = call_arrays(lib, oracle)
    assert rc == 0
    c = [D() for _ in range(5)]
    assert lib.galgetcoefficients_(*[C.byref(v) for v in c]) == 0
    for v, n in zip(c, ("c2", "c3", "c4", "c5", "cG")):
        assert abs(v.value - oracle.get("gal_" + n)) <= 1e-14 * abs(oracle.get("gal_" + n))
    pa, ph, px, n = PD(), PD(), PD(), C.c_int()
    assert lib.galgettable_(C.byref(pa), C.byref(ph), C.byref(px), C.byref(n)) == 0 and n.value == 30000
    a = np.ctypeslib.as_array(pa, (n.value,))
    h = np.ctypeslib.as_array(ph, (n.value,))
    x = np.ctypeslib.as_array(px, (n.value,))
    np.testing.assert_allclose(a, oracle.arr("gal_a")[:30000], rtol=1e-15)
    # two independently written adaptive integrators at rtol 1e-16 over ten decades of a
    np.testing.assert_allclose(h, oracle.arr("gal_h")[:30000], rtol=1e-11)
    np.testing.assert_allclose(x, oracle.arr("gal_x")[:30000], rtol=1e-10)
    for av in (1e-9, 3.3e-7, 1e-4, 0.0123, 0.5, 0.999, 1.0, 1.05):
        assert abs(lib.GetH_(ref(av)) / oracle.L.orc_gal_H(oracle.h, av) - 1) < 1e-10
        assert abs(lib.GetX_(ref(av)) / oracle.L.orc_gal_X(oracle.h, av) - 1) < 1e-9
    assert abs(lib.GetH_(ref(1.0)) - 1) < 1e-12 and abs(lib.GetX_(ref(1.0)) - 1) < 1e-12
    assert np.isnan(lib.GetH_(ref(1.2))) and np.isnan(lib.GetX_(ref(-0.1)))  # the reference exit()s here
    ct = lib.ct_(ref(0.7), gm, nm, C.byref(C.c_int(1)))
    assert 0.0 < ct < 2.0  # tensor sound speed: real (Laplace-stable) at a valid point


def test_closed_forms_against_oracle(lib, oracle):
    rc, gm, nm = call_arrays(lib, oracle)
    assert rc == 0
    rng = np.random.default_rng(5)
    one = C.byref(C.c_int(1))
    lib.GetdHdX_.argtypes = [PD, PD, PD, PD, PD, C.c_void_p, C.c_void_p, C.c_void_p]  # double& = pointer in the ABI
    for a in (2e-6, 1e-4, 3e-3, 0.08, 0.4, 0.9, 1.0):
        inp = np.array([a, 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal(), 1e-4 * rng.standard_normal(),
                        1e-2 * rng.standard_normal(), 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal(),
                        10 ** rng.uniform(-4, -0.5), 1e-4 * rng.standard_normal(), 1e-5 * rng.standard_normal(),
                        abs(rng.standard_normal()) * 1e-6 / a ** 2, abs(rng.standard_normal()) * 3e-7 / a ** 2])
        out = np.zeros(9)
        oracle.L.orc_gal_terms(oracle.h, inp.ctypes.data, out.ctypes.data)
        _, dgrho, dgq, dgpi, eta, dphi, dphip, k, dfp, pidot, grho, gpres = inp
        x = lib.GetX_(ref(a))
        hc = a * lib.GetH_(ref(a))
        dh, dx = D(), D()
        lib.GetdHdX_(ref(a), ref(hc), ref(x), C.byref(dh), C.byref(dx), gm, nm, one)
        got = [dh.value, dx.value,
               lib.grhogal_(ref(a), ref(hc), ref(x)),
               lib.gpresgal_(ref(a), ref(hc), ref(x), ref(dh.value), ref(dx.value)),
               lib.Chigal_(ref(a), ref(hc), ref(x), ref(dgrho), ref(eta), ref(dphi), ref(dphip), ref(k)),
               lib.qgal_(ref(a), ref(hc), ref(x), ref(dgq), ref(eta), ref(dphi), ref(dphip), ref(k)),
               lib.Pigal_(ref(a), ref(hc), ref(x), ref(dh.value), ref(dx.value), ref(dgrho), ref(dgq), ref(dgpi), ref(eta),
                          ref(dphi), ref(k)),
               lib.dphisecond_(ref(a), ref(hc), ref(x), ref(dh.value), ref(dx.value), ref(dgrho), ref(dgq), ref(eta), ref(dphi),
                               ref(dphip), ref(k), ref(dfp)),
               lib.pigalprime_(ref(a), ref(hc), ref(x), ref(dh.value), ref(dx.value), ref(dgrho), ref(dgq), ref(dgpi),
                               ref(pidot), ref(eta), ref(dphi), ref(dphip), ref(k), ref(grho), ref(gpres), gm, nm, one)]
        for name, g_, o_ in zip("dh dx grhogal gpresgal Chigal qgal Pigal dphisecond pigalprime".split(), got, out):
            assert abs(g_ - o_) <= 2e-7 * abs(o_) + 1e-30, (a, name, g_, o_)
    # the hard switch-off below a = 9.99999e-7 (camb/galileon.cc:822 ...)
    a = 5e-7
    assert lib.grhogal_(ref(a), ref(1.0), ref(1.0)) == 0.0


def test_unstable_point_is_rejected_with_a_status(lib, oracle):
    # the uncommented galileon.ini point fails the tensor Laplace condition at a = 1
    rc, _, _ = call_arrays(lib, oracle, c2=-14.539, c3=-5.73, c4=-1.2, cG=0.4)
    assert rc == 7
    assert np.isnan(lib.GetH_(ref(0.5)))  # no stale table after a failed arrays_
    rc, _, _ = call_arrays(lib, oracle)
    assert rc == 0
    lib.freegal_()
    assert np.isnan(lib.GetH_(ref(0.5)))


def test_theta_against_independent_quadrature(lib, oracle):
    """theta_ (camb/theta.cc:296): r_s(a*)/D_A(a*) from its own background run and exact spline integrals, against
    scipy quadrature of the same integrands over the table arrays_ produced.  100 theta of the config-2 point (a chain
    best fit) must also land near the measured acoustic scale."""
    from scipy import integrate
    lib.theta_.restype = D
    rc, gm, nm = call_arrays(lib, oracle)
    assert rc == 0
    g = oracle.get
    h0_kms = g("gal_h0") * 2.99792458e8 / 1000
    ombh2, astar = GAL["ombh2"], 1 / 1090.0
    th = lib.theta_(ref(g("gal_om")), ref(g("gal_orad")), ref(ombh2), ref(h0_kms), ref(astar), ref(g("gal_c2")), ref(g("gal_c3")),
                    ref(g("gal_c4")), ref(g("gal_cG")), gm, nm, C.byref(C.c_int(1)))
    H = lambda a: lib.GetH_(ref(a))
    rs, _ = integrate.quad(lambda a: 1 / (a * a * H(a) * np.sqrt(3 * (1 + 3e4 * a * ombh2))), 1e-8, astar, epsrel=1e-10, limit=400,
                           points=[1e-7, 1e-6, 1e-5, 1e-4])
    da, _ = integrate.quad(lambda a: 1 / (a * a * H(a)), astar, 1.0, epsrel=1e-10, limit=400, points=[0.01, 0.1])
    assert abs(th / (rs / da) - 1) < 1e-7
    assert 1.03 < 100 * th < 1.05
    assert abs(lib.GetH_(ref(1.0)) - 1) < 1e-12  # arrays_' tables untouched by theta_
    # rejected background -> -1 (camb/theta.cc:380)
    bad = lib.theta_(ref(g("gal_om")), ref(g("gal_orad")), ref(ombh2), ref(h0_kms), ref(astar), ref(-14.539), ref(-5.73), ref(-1.2),
                     ref(0.4), gm, nm, C.byref(C.c_int(1)))
    assert bad == -1.0
