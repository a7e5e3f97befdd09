"""ctypes front-end to the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liborc.so")
NATIVE_LIB = os.path.join(ROOT, "oracle", "_build", "native", "liborc.so")
TEMPLATE_F64 = os.path.join(ROOT, "tests", "golden", "highL_template.f64")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8"])


def use_native_build():
    """bench.py's CPU-baseline legs: switch this process to the -O3 -march=native build of the oracle (BASELINE.md
    section 2), compiled on this machine; returns False (and keeps the strict build) if that build fails or if the
    library was already loaded."""
    global LIB
    if _lib is not None:
        return LIB == NATIVE_LIB
    try:
        # -march=native objects are only valid on the CPU model that built them (the tree travels between machines)
        cpu = ""
        try:
            cpu = next(ln for ln in open("/proc/cpuinfo") if ln.startswith("model name")).strip()
        except Exception:  # noqa: BLE001
            pass
        stamp = os.path.join(os.path.dirname(NATIVE_LIB), ".cpu")
        same = os.path.exists(stamp) and open(stamp).read() == cpu
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8", "native"] + ([] if same else ["-B"]))
        open(stamp, "w").write(cpu)
    except Exception:  # noqa: BLE001
        return False
    LIB = NATIVE_LIB
    return True


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_new.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_set_physical.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.orc_load_template.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_run.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        for s in ("background", "initvars", "sources", "spline_k", "bessels", "los", "cls", "lensing", "qgrid", "transfer"):
            getattr(L, "orc_stage_" + s).argtypes = [C.c_void_p]
        L.orc_alloc_sources.argtypes = [C.c_void_p]
        L.orc_set_transfer_redshifts.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_transfer.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_source_k.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_get.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get.restype = C.c_double
        L.orc_get_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.orc_get_array.restype = C.c_long
        L.orc_put_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.orc_put_array.restype = C.c_long
        L.orc_bjl.argtypes = [C.c_int, C.c_double]
        L.orc_bjl.restype = C.c_double
        L.orc_spline.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_splder.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_rombint_kat.argtypes = [C.c_double] * 3
        L.orc_rombint_kat.restype = C.c_double
        L.orc_dverk_kat.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_void_p]
        L.orc_ranges_kat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_ranges_indexof.argtypes = [C.c_void_p, C.c_int, C.c_double]
        for f in ("orc_dtauda", "orc_gal_H", "orc_gal_X"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double]
            getattr(L, f).restype = C.c_double
        L.orc_gal_tracker.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_thermo.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_nu_background.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_gal_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# The five BASELINE.json configurations as parameter dictionaries.
LCDM = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.00064, H0=70.0)
GALILEON = dict(ombh2=0.02258, omch2=0.1109, omnuh2=0.00064, H0=71.0, use_galileon=1, c2=-14.539, c3=-5.73, c4=-1.2,
                cG=0.4)


class Oracle:
    """One oracle model instance.  kw: physical densities + any orc_set() key."""

    def __init__(self, template=True, lmax=2500, **kw):
        self.L = lib()
        self.h = self.L.orc_new()
        p = dict(kw)
        self.L.orc_set_physical(self.h, p.pop("ombh2"), p.pop("omch2"), p.pop("omnuh2"), p.pop("H0"))
        self.set(Max_l=lmax, Max_eta_k=2.0 * lmax)
        self.set(**p)
        if template:
            rc = self.L.orc_load_template(self.h, TEMPLATE_F64.encode())
            if rc:
                raise RuntimeError("template fixture missing: " + TEMPLATE_F64)
        else:
            self.set(use_spline_template=0)

    def set_point(self, **kw):
        """New parameter point on an existing model object (keeps its cached Bessel tables, camb/bessels.f90:40-41)."""
        p = dict(kw)
        self.L.orc_set_physical(self.h, p.pop("ombh2"), p.pop("omch2"), p.pop("omnuh2"), p.pop("H0"))
        self.set(**p)

    def set(self, **kw):
        for k, v in kw.items():
            if self.L.orc_set(self.h, k.encode(), float(v)):
                raise KeyError(k)

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError("oracle error %d: %s" % (rc, self.L.orc_error(self.h).decode()))

    def run(self, reuse_bessels=False):
        st = (C.c_double * 8)()
        self._chk(self.L.orc_run(self.h, st, int(reuse_bessels)))
        return dict(zip(("params_set", "initvars", "sources", "spline_k", "bessels", "los", "cls", "total"), list(st)))

    def stage(self, *names):
        for n in names:
            self._chk(getattr(self.L, "orc_stage_" + n)(self.h))

    def get(self, name):
        return self.L.orc_get(self.h, name.encode())

    def geti(self, name):
        return int(round(self.get(name)))

    def arr(self, name):
        n = self.L.orc_get_array(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=np.float64)
        self.L.orc_get_array(self.h, name.encode(), out.ctypes.data, n)
        return out

    def put(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        rc = self.L.orc_put_array(self.h, name.encode(), a.ctypes.data, a.size)
        if rc < 0:
            raise ValueError("put %s failed (%d)" % (name, rc))

    # shaped views (Fortran order of the reference arrays)
    def Src(self, name="Src"):
        nk, S, nt = self.geti("nk"), self.geti("SourceNum"), self.geti("nt")
        return self.arr(name).reshape(nt, S, nk)  # [t][s][k]  == Fortran Src(k,s,t)

    def Delta(self):
        nq, S, nl = self.geti("nq"), self.geti("SourceNum"), self.geti("nl")
        return self.arr("Delta_p_l_k").reshape(nq, nl, S)  # [q][l][s] == Fortran Delta_p_l_k(s,l,q)

    def Cl(self):
        nL = self.geti("maximum_l") - 1
        return self.arr("Cl").reshape(self.geti("C_last"), nL)  # [type][l-2], dimensionless l(l+1)C_l/2pi

    def Cl_lensed(self):
        nLl = self.geti("lmax_lensed") - 1
        return self.arr("Cl_lensed").reshape(4, nLl)  # [TT,EE,BB,TE][l-2]

    def iCl(self):
        return self.arr("iCl").reshape(self.geti("C_last"), self.geti("nl"))

    # matter transfer functions (camb/cmbmain.f90:508-632, :1151-1201; camb/modules.f90:1876-1915)
    def want_transfer(self, redshifts=(0.0,), kmax=0.9, k_per_logint=0):
        """Transfer outputs at the given redshifts (descending, like CP%Transfer%redshifts); kmax in 1/Mpc."""
        z = np.ascontiguousarray(redshifts, dtype=np.float64)
        if self.L.orc_set_transfer_redshifts(self.h, z.ctypes.data, len(z)):
            raise ValueError("too many transfer redshifts")
        self.set(WantTransfer=1, transfer_kmax=kmax, transfer_k_per_logint=k_per_logint)

    def transfer(self):
        """dict: TransferData [nz][nq][13] (= Fortran TransferData(13, nq, nz), FP64), q_trans, tautf, sigma8."""
        dims = (C.c_int * 3)()
        self.L.orc_get_transfer(self.h, dims, None, None, None, None)
        ne, nq, nz = list(dims)
        T, q, t, s8 = np.zeros((nz, nq, ne)), np.zeros(nq), np.zeros(nz), np.zeros(nz)
        self.L.orc_get_transfer(self.h, dims, T.ctypes.data, q.ctypes.data, t.ctypes.data, s8.ctypes.data)
        return dict(TransferData=T, q_trans=q, tautf=t, sigma8=s8)
