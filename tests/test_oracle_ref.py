"""Pins the oracle's Galileon terms against the REFERENCE'S OWN OBJECT CODE.

oracle/_ref/libgalref.so is camb/galileon.cc of the reference, unmodified, compiled where it lies by oracle/Makefile
(target `ref`) against stand-in GSL headers (oracle/ref/gsl_shim) and with the three MassiveNu call-backs served from
the oracle's neutrino tables (oracle/ref/ref_callbacks.cpp).  The closed forms GetdHdX_, ct_, grhogal_, gpresgal_,
Chigal_, qgal_, Pigal_, dphisecond_, pigalprime_ (camb/galileon.cc:746-1174) involve no GSL at all, so these checks
compare the oracle's restatement with the reference's compiled arithmetic directly.  arrays_ (the background) runs
through the shim's RKF45 / natural spline, i.e. reference code over a restated GSL.

/root/reference is not present on the GPU box; the prebuilt .so travels.  Without either, the tests skip.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_py import GALILEON, ROOT, Oracle, lib

REF = os.path.join(ROOT, "oracle", "_ref", "libgalref.so")
PD = C.POINTER(C.c_double)


def _ref():
    lib()  # liborc.so first (libgalref links against it)
    if os.path.exists("/root/reference/camb/galileon.cc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libgalref.so not built and /root/reference absent")
    R = C.CDLL(REF, mode=C.RTLD_GLOBAL)
    R.ref_set_nu_model.argtypes = [C.c_void_p]
    R.ref_set_globals.argtypes = [C.c_double] * 8
    R.ref_get_globals.argtypes = [C.c_void_p]
    R.ref_set_background.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    for f in ("GetX_", "GetH_"):
        getattr(R, f).argtypes = [PD]
        getattr(R, f).restype = C.c_double
    R.arrays_.argtypes = [PD] * 9 + [C.POINTER(C.c_int)]
    R.arrays_.restype = C.c_int
    R.ct_.argtypes = [PD, PD, PD, C.POINTER(C.c_int)]
    R.ct_.restype = C.c_double
    R.ref_GetdHdX.argtypes = [PD] * 7 + [C.POINTER(C.c_int)]
    for f, n in (("grhogal_", 3), ("gpresgal_", 5), ("Chigal_", 8), ("qgal_", 8), ("Pigal_", 11), ("dphisecond_", 12)):
        getattr(R, f).argtypes = [PD] * n
        getattr(R, f).restype = C.c_double
    R.pigalprime_.argtypes = [PD] * 17 + [C.POINTER(C.c_int)]
    R.pigalprime_.restype = C.c_double
    return R


# the config-2 parameter point (the in-file test point of camb/galileon.cc:1411-1418)
CONFIG2 = dict(GALILEON, c2=-8.438179, c3=-3.366555, c4=-1.03411, cG=0.001188243, H0=78.07227, ombh2=0.0221743,
               omch2=0.1177806)


def _oracle(n_eig):
    o = Oracle(**CONFIG2)
    if n_eig == 0:
        o.set(Nu_mass_eigenstates=0, Num_Nu_massive=0, Num_Nu_massless=3.046)
        o.L.orc_set_physical(o.h, CONFIG2["ombh2"], CONFIG2["omch2"], 0.0, CONFIG2["H0"])
    elif n_eig == 3:  # CosmoMC's degenerate hierarchy (source/Calculator_CAMB.f90:139,866-867)
        o.set(Nu_mass_eigenstates=3, Num_Nu_massive=3, Num_Nu_massless=0.046)
        for i in (1, 2, 3):
            o.set(**{"Nu_mass_numbers%d" % i: 1, "Nu_mass_fractions%d" % i: 1.0 / 3.0})
    o.stage("background")
    return o


def _d(v):
    return C.byref(C.c_double(v))


def _nu_arrays(o, n_eig):
    gm = (C.c_double * 5)(*([o.get("grhormass%d" % i) for i in range(1, n_eig + 1)] + [0.0] * (5 - n_eig)))
    nm = (C.c_double * 5)(*([o.get("nu_masses%d" % i) for i in range(1, n_eig + 1)] + [0.0] * (5 - n_eig)))
    return gm, nm, C.c_int(n_eig)


def _close(a, b, rtol, scale=None, what=""):
    a, b = float(a), float(b)
    s = max(abs(a), abs(b)) if scale is None else scale
    assert abs(a - b) <= rtol * s + 1e-300, "%s: oracle %.17g vs reference %.17g" % (what, a, b)


@pytest.mark.parametrize("n_eig", [0, 1, 3])
def test_closed_forms_against_reference_object_code(n_eig):
    """GetdHdX_ ... pigalprime_, ct_ over a grid of (a, k, perturbations), incl. the a < 9.99999e-7 zeroing."""
    R = _ref()
    o = _oracle(n_eig)
    R.ref_set_nu_model(o.h)
    g = o.get
    R.ref_set_globals(g("gal_om"), g("gal_orad"), g("gal_h0"), g("gal_c2"), g("gal_c3"), g("gal_c4"), g("gal_c5"),
                      g("gal_cG"))
    # the reference's spline globals get the ORACLE's table: both sides then evaluate the closed forms at the same
    # (a, H, x) up to the rounding of two independent natural-spline evaluations
    ga, gh, gx = o.arr("gal_a"), o.arr("gal_h"), o.arr("gal_x")
    R.ref_set_background(ga.ctypes.data, gh.ctypes.data, gx.ctypes.data, ga.size)
    gm, nm, ne = _nu_arrays(o, n_eig)
    rng = np.random.default_rng(7)
    a_grid = [2e-9, 5e-7, 9.99998e-7, 9.99999e-7, 1.0000001e-6, 3e-6, 1e-5, 1e-4, 9.1e-4, 1e-3, 0.01, 0.09, 0.3, 0.5,
              0.77, 0.999, 1.0]
    n_zero = 0
    for a in a_grid:
        for k in (1e-4, 3e-3, 0.05, 0.4):
            pert = rng.normal(size=10) * 10.0 ** rng.uniform(-6, 0, size=10)
            dgrho, dgq, dgpi, eta, dphi, dphip, dfp, pidot = pert[:8]
            grho, gpres = abs(pert[8]) + 1e-4, abs(pert[9]) * 0.3
            inp = np.array([a, dgrho, dgq, dgpi, eta, dphi, dphip, k, dfp, pidot, grho, gpres])
            out = np.zeros(9)
            o.L.orc_gal_terms(o.h, inp.ctypes.data, out.ctypes.data)
            x_o, hb_o = o.L.orc_gal_X(o.h, a), o.L.orc_gal_H(o.h, a)
            x_r, hb_r = R.GetX_(_d(a)), R.GetH_(_d(a))
            _close(x_o, x_r, 1e-14, what="GetX a=%g" % a)
            _close(hb_o, hb_r, 1e-14, what="GetH a=%g" % a)
            hc = a * hb_r
            dh, dx = C.c_double(), C.c_double()
            R.ref_GetdHdX(_d(a), _d(hc), _d(x_r), C.byref(dh), C.byref(dx), gm, nm, C.byref(ne))
            # dh, dx are quotients of near-cancelling polynomials; compare on the scale of their operands
            _close(out[0], dh.value, 1e-11, scale=max(abs(dh.value), hb_r), what="dh a=%g" % a)
            _close(out[1], dx.value, 1e-11, scale=max(abs(dx.value), abs(x_r)), what="dx a=%g" % a)
            A = [_d(a), _d(hc), _d(x_r)]
            D = [_d(dh.value), _d(dx.value)]
            ref = [
                R.grhogal_(*A),
                R.gpresgal_(*A, *D),
                R.Chigal_(*A, _d(dgrho), _d(eta), _d(dphi), _d(dphip), _d(k)),
                R.qgal_(*A, _d(dgq), _d(eta), _d(dphi), _d(dphip), _d(k)),
                R.Pigal_(*A, *D, _d(dgrho), _d(dgq), _d(dgpi), _d(eta), _d(dphi), _d(k)),
                R.dphisecond_(*A, *D, _d(dgrho), _d(dgq), _d(eta), _d(dphi), _d(dphip), _d(k), _d(dfp)),
                R.pigalprime_(*A, *D, _d(dgrho), _d(dgq), _d(dgpi), _d(pidot), _d(eta), _d(dphi), _d(dphip), _d(k),
                              _d(grho), _d(gpres), gm, nm, C.byref(ne)),
            ]
            names = "grhogal gpresgal Chigal qgal Pigal dphisecond pigalprime".split()
            if a < 9.99999e-7:
                assert all(r == 0.0 for r in ref), "reference terms must be hard-zeroed below a = 9.99999e-7"
                assert all(v == 0.0 for v in out[2:]), "oracle terms must be hard-zeroed below a = 9.99999e-7"
                n_zero += 1
                continue
            # the oracle is handed (dh, dx) it computed itself: 1e-11 of their scale propagates linearly
            for nme, vo, vr in zip(names, out[2:], ref):
                _close(vo, vr, 2e-10 if nme in ("gpresgal", "Pigal", "dphisecond", "pigalprime") else 1e-13,
                       what="%s a=%g k=%g" % (nme, a, k))
    assert n_zero == 3 * 4
    # ct_ (tensor speed, camb/galileon.cc:789-807): the product's and the oracle's stability checks use the same form
    for a in (1e-3, 0.1, 1.0):
        ct_r = R.ct_(_d(a), gm, nm, C.byref(ne))
        assert np.isfinite(ct_r) and ct_r > 0


def test_closed_forms_bitwise_when_fed_the_same_dh_dx():
    """With (a, H, x, dh, dx) taken from ONE side the remaining closed forms are pure polynomial arithmetic:
    the restatement must reproduce the reference's object code to the last few ulps."""
    R = _ref()
    o = _oracle(1)
    R.ref_set_nu_model(o.h)
    g = o.get
    R.ref_set_globals(g("gal_om"), g("gal_orad"), g("gal_h0"), g("gal_c2"), g("gal_c3"), g("gal_c4"), g("gal_c5"),
                      g("gal_cG"))
    ga, gh, gx = o.arr("gal_a"), o.arr("gal_h"), o.arr("gal_x")
    R.ref_set_background(ga.ctypes.data, gh.ctypes.data, gx.ctypes.data, ga.size)
    gm, nm, ne = _nu_arrays(o, 1)
    worst = 0.0
    for a in (2e-6, 1e-4, 1e-2, 0.2, 0.6, 1.0):
        inp = np.array([a, 3e-3, -2e-4, 1e-5, 0.4, 2e-3, -7e-3, 0.02, 5e-4, 3e-6, 0.7, 0.1])
        out = np.zeros(9)
        o.L.orc_gal_terms(o.h, inp.ctypes.data, out.ctypes.data)
        x, hc = o.L.orc_gal_X(o.h, a), a * o.L.orc_gal_H(o.h, a)
        A = [_d(a), _d(hc), _d(x)]
        D = [_d(out[0]), _d(out[1])]  # the oracle's own dh, dx
        p = [_d(v) for v in inp]
        ref = [R.grhogal_(*A), R.gpresgal_(*A, *D), R.Chigal_(*A, p[1], p[4], p[5], p[6], p[7]),
               R.qgal_(*A, p[2], p[4], p[5], p[6], p[7]), R.Pigal_(*A, *D, p[1], p[2], p[3], p[4], p[5], p[7]),
               R.dphisecond_(*A, *D, p[1], p[2], p[4], p[5], p[6], p[7], p[8]),
               R.pigalprime_(*A, *D, p[1], p[2], p[3], p[9], p[4], p[5], p[6], p[7], p[10], p[11], gm, nm, C.byref(ne))]
        for vo, vr in zip(out[2:], ref):
            worst = max(worst, abs(vo - vr) / max(abs(vr), 1e-300))
    assert worst < 1e-13, worst


@pytest.mark.parametrize("n_eig", [1, 3])
def test_background_arrays_against_reference(n_eig):
    """arrays_ of the reference (its own RHS calcValOmC2C3C4C5CGC0, stability checks calcPertOmC2C3C4C5CGC0, Friedmann
    closure, c5 closure, spline set-up) over the stand-in GSL, against the oracle's gal_arrays."""
    R = _ref()
    o = _oracle(n_eig)
    R.ref_set_nu_model(o.h)
    g = o.get
    gm, nm, ne = _nu_arrays(o, n_eig)
    rc = R.arrays_(_d(g("gal_orad")), _d(g("gal_om")), _d(g("gal_h0")), _d(g("gal_c2")), _d(CONFIG2["c3"]),
                   _d(CONFIG2["c4"]), _d(g("gal_cG")), gm, nm, C.byref(ne))
    assert rc == 0
    gl = np.zeros(9)
    R.ref_get_globals(gl.ctypes.data)
    _close(gl[6], g("gal_c5"), 1e-14, what="c5 closure")
    _close(gl[4], g("gal_c3"), 0, what="c3")
    worst_h = worst_x = 0.0
    for a in np.geomspace(1.0000001e-10, 1.0, 400):
        hr, xr = R.GetH_(_d(a)), R.GetX_(_d(a))
        ho, xo = o.L.orc_gal_H(o.h, a), o.L.orc_gal_X(o.h, a)
        worst_h = max(worst_h, abs(hr - ho) / abs(hr))
        worst_x = max(worst_x, abs(xr - xo) / max(abs(xr), 1e-30))
    # measured: bit-identical (0.0) -- the reference's own right-hand side and checks under two independently
    # written RKF45 / natural-spline restatements of GSL's documented algorithms take the same steps
    assert worst_h < 1e-13, worst_h
    assert worst_x < 1e-13, worst_x
