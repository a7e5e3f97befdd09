"""ctypes binding of libgalcamb_b200.so + a thin host class following CAMB's own stage order.

Reference call sequence being mirrored (camb/cmbmain.f90:131-291):
    InitVars -> SetkValuesForSources -> [loop DoSourcek] -> InitSourceInterpolation -> InitSpherBessels ->
    SetkValuesForInt -> [loop SourceToTransfers] -> ClTransferToCl (CalcScalCls, InterpolateCls)
The bracketed loops and the two routines after them are the GPU stages; everything before is the caller's
(the Fortran host in production, a table fixture or the test oracle here) and arrives as ModelTables.
"""
import ctypes as C
import os

import numpy as np

MAX_NU, MAX_NQ, MAX_REGIONS, MAX_TRANSFER_Z = 5, 5, 16, 16
_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # GALCAMB_LIB selects an alternative build of the same library (kernel experiments); default = in-tree build
    return os.environ.get("GALCAMB_LIB") or os.path.join(_HERE, "libgalcamb_b200.so")


class Region(C.Structure):
    _fields_ = [("Low", C.c_double), ("High", C.c_double), ("delta", C.c_double), ("start_index", C.c_int),
                ("IsLog", C.c_int)]


_D, _I, _P = C.c_double, C.c_int, C.POINTER(C.c_double)


class CModel(C.Structure):
    """struct galcamb_model of include/galcamb.h (field order must match)."""
    _fields_ = (
        [(n, _D) for n in ("grhom grhog grhor grhob grhoc grhov grhornomass grhok adotrad tau0 taurst taurend tau_maxvis "
                           "tight_tau matter_verydom_tau epsw max_etak_scalar AccuracyBoost lAccuracyBoost").split()]
        + [(n, _I) for n in ("use_galileon DoLateRadTruncation HighAccuracyDefault AccuratePolarization "
                             "AccurateReionization WantLateTime MassiveNuMethod Nu_mass_eigenstates Num_Nu_massive "
                             "nqmax").split()]
        + [("grhormass", _D * MAX_NU), ("nu_masses", _D * MAX_NU), ("nu_q", _D * MAX_NQ), ("nu_int_kernel", _D * MAX_NQ),
           ("nu_tau_notmassless", (_D * MAX_NU) * MAX_NQ), ("nu_tau_nonrelativistic", _D * MAX_NU),
           ("nu_tau_massive", _D * MAX_NU), ("dlnam", _D)]
        + [(n, _P) for n in "nu_r1 nu_p1 nu_dr1 nu_dp1 nu_ddr1 nu_ddp1".split()]
        + [("nthermo", _I), ("tauminn", _D), ("dlntau", _D)]
        + [(n, _P) for n in "cs2 dcs2 dotmu ddotmu dddotmu".split()]
        + [("nt", _I)]
        + [(n, _P) for n in "tau dtau vis dvis ddvis expmmu opac dopac".split()]
        + [("n_tau_regions", _I), ("tau_regions", Region * MAX_REGIONS)]
        + [(n, _D) for n in "c2 c3 c4 c5 cG h0 orad".split()]
        + [("ngal", _I)]
        + [(n, _P) for n in "gal_a gal_h gal_x".split()]
        + [("nk", _I), ("k", _P), ("nq", _I), ("q", _P), ("dq", _P)]
        + [(n, _D) for n in "ScalarPowerAmp an n_run n_runrun k_0_scalar".split()]
        + [("lmax", _I)]
        + [("WantTransfer", _I), ("transfer_num_redshifts", _I), ("tautf", _D * MAX_TRANSFER_Z), ("nk_transfer", _I),
           ("k_transfer", _P), ("H0", _D), ("ALens", _D), ("ALens_Fiducial", _D)]
    )


class CParams(C.Structure):
    """struct galcamb_params of include/galcamb.h (field order must match): the CAMBparams subset of the scalar path."""
    _fields_ = (
        [(n, _D) for n in "H0 omegab omegac omegav omegan TCMB YHe Num_Nu_massless".split()]
        + [(n, _I) for n in "Num_Nu_massive Nu_mass_eigenstates share_delta_neff".split()]
        + [("Nu_mass_degeneracies", _D * MAX_NU), ("Nu_mass_fractions", _D * MAX_NU), ("Nu_mass_numbers", _I * MAX_NU)]
        + [("use_galileon", _I)]
        + [(n, _D) for n in "c2 c3 c4 cG ScalarPowerAmp an n_run n_runrun k_0_scalar".split()]
        + [("Reionization", _I), ("use_optical_depth", _I)]
        + [(n, _D) for n in ("optical_depth reion_redshift delta_redshift reion_fraction helium_redshift "
                             "helium_delta_redshift helium_redshiftstart RECFAST_fudge RECFAST_fudge_He").split()]
        + [("RECFAST_Heswitch", _I), ("RECFAST_Hswitch", _I), ("Max_l", _I), ("Max_eta_k", _D)]
        + [(n, _I) for n in ("DoLensing AccuratePolarization AccurateReionization MassiveNuMethod DoLateRadTruncation "
                             "HighAccuracyDefault").split()]
        + [(n, _D) for n in "AccuracyBoost lSampleBoost lAccuracyBoost".split()]
        + [("use_spline_template", _I)]
        + [(n, _I) for n in "WantTransfer transfer_high_precision transfer_k_per_logint transfer_num_redshifts".split()]
        + [("transfer_kmax", _D), ("transfer_redshifts", _D * MAX_TRANSFER_Z), ("ALens", _D), ("ALens_Fiducial", _D)]
    )


def make_params(ombh2=None, omch2=None, omnuh2=None, **kw):
    """CParams from CAMB-style keywords: the defaults of galcamb_default_params_, then physical densities (ombh2, omch2,
    omnuh2 -> omegab, omegac, omegan with omegav closing the universe, as camb/inidriver.F90:108-121 does) and any
    field of galcamb_params by name.  Arrays (Nu_mass_*) take sequences."""
    p = CParams()
    load_library().galcamb_default_params_(C.byref(p))
    H0 = kw.get("H0", p.H0)
    h2 = (H0 / 100.0) ** 2
    if ombh2 is not None:
        p.omegab = ombh2 / h2
    if omch2 is not None:
        p.omegac = omch2 / h2
    if omnuh2 is not None:
        p.omegan = omnuh2 / h2
    names = {n for n, _ in CParams._fields_}
    for k, v in kw.items():
        if k not in names:
            raise KeyError("galcamb_params has no field %r" % k)
        if (k.startswith("Nu_mass_") and k != "Nu_mass_eigenstates") or k == "transfer_redshifts":
            arr = getattr(p, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(p, k, v)
    if "omegav" not in kw:
        p.omegav = 1.0 - p.omegab - p.omegac - p.omegan
    return p


SCALARS = [n for n, t in CModel._fields_ if t in (_D, _I)]
ARRAYS = [n for n, t in CModel._fields_ if t is _P]
SMALL = ["grhormass", "nu_masses", "nu_q", "nu_int_kernel", "nu_tau_notmassless", "nu_tau_nonrelativistic", "nu_tau_massive"]


class ModelTables:
    """The per-model state the host owns after InitVars, as numpy arrays + scalars (see galcamb_model)."""

    def __init__(self, d):
        self.d = {}
        for k, v in d.items():
            if k in ARRAYS:
                self.d[k] = np.ascontiguousarray(v, dtype=np.float64)
            elif k in ("ls", "tau_regions"):
                self.d[k] = np.asarray(v)
            elif k in SMALL:
                self.d[k] = np.asarray(v, dtype=np.float64)
            else:
                self.d[k] = v
        self._c = None

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        return cls({k: (z[k] if z[k].ndim else z[k].item()) for k in z.files})

    def save(self, path):
        np.savez_compressed(path, **self.d)

    def copy(self, **override):
        d = dict(self.d)
        d.update(override)
        return ModelTables(d)

    @property
    def ls(self):
        return np.asarray(self.d["ls"], dtype=np.int32)

    def nbytes(self):
        return sum(self.d[k].nbytes for k in ARRAYS if k in self.d)

    def cstruct(self):
        if self._c is not None:
            return self._c
        m = CModel()
        d = self.d
        for n in SCALARS:
            if n in ("WantTransfer", "transfer_num_redshifts", "nk_transfer", "H0", "ALens", "ALens_Fiducial"):  # absent from older fixtures
                setattr(m, n, d.get(n, 0))
            elif n != "n_tau_regions":
                setattr(m, n, d[n])
        tf = np.zeros(MAX_TRANSFER_Z)
        a = np.atleast_1d(d.get("tautf", []))
        tf[: len(a)] = a[:MAX_TRANSFER_Z]
        m.tautf = (_D * MAX_TRANSFER_Z)(*tf)
        for n in ARRAYS:
            a = d.get(n)
            setattr(m, n, a.ctypes.data_as(_P) if a is not None and a.size else None)
        for n in ("grhormass", "nu_masses", "nu_tau_nonrelativistic", "nu_tau_massive"):
            v = np.zeros(MAX_NU)
            a = np.atleast_1d(d.get(n, []))
            v[: len(a)] = a[:MAX_NU]
            setattr(m, n, (_D * MAX_NU)(*v))
        for n in ("nu_q", "nu_int_kernel"):
            v = np.zeros(MAX_NQ)
            a = np.atleast_1d(d.get(n, []))
            v[: len(a)] = a[:MAX_NQ]
            setattr(m, n, (_D * MAX_NQ)(*v))
        t = np.zeros((MAX_NQ, MAX_NU))
        a = np.atleast_2d(d.get("nu_tau_notmassless", np.zeros((0, 0))))
        if a.size:
            t[: a.shape[0], : a.shape[1]] = a
        for i in range(MAX_NQ):
            for j in range(MAX_NU):
                m.nu_tau_notmassless[i][j] = t[i, j]
        reg = np.asarray(d["tau_regions"], dtype=np.float64).reshape(-1, 5)
        m.n_tau_regions = len(reg)
        for i, r in enumerate(reg):
            m.tau_regions[i] = Region(r[0], r[1], r[2], int(r[3]), int(r[4]))
        self._c = m
        return m


_lib = None


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p) and not os.environ.get("GALCAMB_LIB"):
            try:  # the .so is a build artefact (git-ignored): compile it in place when nvcc is around
                from . import build as _build
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError("libgalcamb_b200.so is not built and could not be built (%s); there is no CPU "
                                   "fallback" % e)
        if not os.path.exists(p):
            raise RuntimeError("libgalcamb_b200.so not found at %s; there is no CPU fallback" % p)
        L = C.CDLL(p)
        L.galcamb_init_.argtypes = [C.POINTER(_I)]
        L.galcamb_last_error_.restype = C.c_char_p
        L.galcamb_set_bessels_.argtypes = [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_D)]
        L.galcamb_get_bessels_.argtypes = [C.POINTER(_I), C.c_void_p, C.c_void_p, C.c_void_p]
        L.galcamb_set_template_.argtypes = [C.c_void_p, C.POINTER(_I)]
        L.galcamb_set_lensing_template_.argtypes = [C.c_void_p, C.POINTER(_I)]
        L.galcamb_get_lensed_cls_.argtypes = [C.POINTER(_I), C.POINTER(_I), C.c_void_p]
        L.galcamb_batch_begin_.argtypes = [C.POINTER(_I)]
        L.galcamb_set_model_.argtypes = [C.POINTER(_I), C.POINTER(CModel)]
        L.galcamb_get_sources_.argtypes = [C.POINTER(_I), C.c_void_p]
        L.galcamb_put_sources_.argtypes = [C.POINTER(_I), C.c_void_p]
        L.galcamb_get_transfers_.argtypes = [C.POINTER(_I), C.c_void_p]
        L.galcamb_get_cls_.argtypes = [C.POINTER(_I), C.c_void_p, C.c_void_p]
        L.galcamb_put_cls_.argtypes = [C.POINTER(_I), C.c_void_p]
        L.galcamb_get_status_.argtypes = [C.POINTER(_I), C.POINTER(_I)]
        L.galcamb_get_transfer_.argtypes = [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.c_void_p, C.c_void_p, C.c_void_p]
        L.galcamb_batch_cls_.argtypes = [C.POINTER(CModel), C.POINTER(_I), C.c_void_p]
        L.galcamb_batch_lensed_cls_.argtypes = [C.POINTER(CModel), C.POINTER(_I), C.c_void_p, C.POINTER(_I), C.c_void_p]
        L.galcamb_set_primordial_.argtypes = [C.POINTER(_I)] + [C.POINTER(_D)] * 5
        L.galcamb_transfers_to_powers_.argtypes = [C.POINTER(_I)]
        L.galcamb_get_timings_.argtypes = [C.c_void_p]
        L.galcamb_get_counters_.argtypes = [C.c_void_p]
        L.galcamb_get_evolve_stats_.argtypes = [C.c_void_p]
        L.galcamb_default_params_.argtypes = [C.POINTER(CParams)]
        L.galcamb_default_params_.restype = None
        L.galcamb_prestage_.argtypes = [C.POINTER(CParams), C.POINTER(_I), C.POINTER(_I)]
        L.galcamb_prestage_status_.argtypes = [C.POINTER(_I), C.POINTER(_I)]
        L.galcamb_prestage_error_.restype = C.c_char_p
        L.galcamb_prestage_model_.argtypes = [C.POINTER(_I), C.POINTER(CModel)]
        L.galcamb_prestage_lsamples_.argtypes = [C.POINTER(_I), C.POINTER(_I), C.c_void_p, C.POINTER(_I)]
        L.galcamb_prestage_free_.restype = None
        L.galcamb_params_to_cls_.argtypes = [C.POINTER(CParams), C.POINTER(_I), C.POINTER(_I), C.c_void_p]
        L.galcamb_prestage_async_.argtypes = [C.POINTER(CParams), C.POINTER(_I), C.POINTER(_I)]
        L.galcamb_prestaged_to_cls_.argtypes = [C.c_void_p]
        L.galcamb_prestaged_upload_.argtypes = []
        L.galcamb_set_host_threads_.argtypes = [C.POINTER(_I)]
        L.galcamb_set_host_threads_.restype = None
        L.galcamb_set_prestage_device_.argtypes = [C.POINTER(_I)]
        L.galcamb_select_device_.argtypes = [C.POINTER(_I)]
        L.galcamb_multi_params_to_cls_.argtypes = [C.POINTER(_I), C.POINTER(CParams), C.POINTER(_I), C.POINTER(_I), C.c_void_p]
        L.galcamb_multi_single_cls_.argtypes = [C.POINTER(_I), C.POINTER(CParams), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class GalCambError(RuntimeError):
    pass


def prestage(params, nthreads=0, copy=True):
    """galcamb_prestage_ for a list of CParams: CAMBParams_Set + InitVars + the k grids + l sampling on the host threads
    (no GPU needed).  Returns (tables, status): tables[i] is a ModelTables (None for a rejected point), status[i] the
    CAMB error id.  copy=False leaves the arrays as views into the library's buffers (valid until the next call)."""
    L = load_library()
    n = len(params)
    arr = (CParams * n)(*params)
    L.galcamb_prestage_(arr, C.byref(_I(n)), C.byref(_I(nthreads)))
    tables, status = [], []
    for i in range(n):
        st = _I(0)
        L.galcamb_prestage_status_(C.byref(_I(i)), C.byref(st))
        status.append(st.value)
        if st.value:
            tables.append(None)
            continue
        m = CModel()
        rc = L.galcamb_prestage_model_(C.byref(_I(i)), C.byref(m))
        assert rc == 0
        nl, kmax = _I(0), _I(0)
        L.galcamb_prestage_lsamples_(C.byref(_I(i)), C.byref(nl), None, C.byref(kmax))
        ls = np.zeros(nl.value, dtype=np.int32)
        L.galcamb_prestage_lsamples_(C.byref(_I(i)), C.byref(nl), ls.ctypes.data, C.byref(kmax))
        d = {}
        for name in SCALARS:
            if name != "n_tau_regions":
                d[name] = getattr(m, name)
        sizes = dict(nu_r1=2000, nu_p1=2000, nu_dr1=2000, nu_dp1=2000, nu_ddr1=2000, nu_ddp1=2000, cs2=m.nthermo,
                     dcs2=m.nthermo, dotmu=m.nthermo, ddotmu=m.nthermo, dddotmu=m.nthermo, tau=m.nt, dtau=m.nt, vis=m.nt,
                     dvis=m.nt, ddvis=m.nt, expmmu=m.nt, opac=m.nt, dopac=m.nt, gal_a=m.ngal, gal_h=m.ngal, gal_x=m.ngal,
                     k=m.nk, q=m.nq, dq=m.nq, k_transfer=m.nk_transfer)
        for name in ARRAYS:
            ptr = getattr(m, name)
            if not ptr or sizes[name] == 0:
                d[name] = np.zeros(0)
            else:
                a = np.ctypeslib.as_array(ptr, shape=(sizes[name],))
                d[name] = a.copy() if copy else a
        ne = m.Nu_mass_eigenstates
        d["grhormass"] = np.array(m.grhormass[:ne])
        d["nu_masses"] = np.array(m.nu_masses[:ne])
        d["nu_tau_nonrelativistic"] = np.array(m.nu_tau_nonrelativistic[:ne])
        d["nu_tau_massive"] = np.array(m.nu_tau_massive[:ne])
        nqm = m.nqmax if ne else 0
        d["nu_q"] = np.array(m.nu_q[:nqm])
        d["nu_int_kernel"] = np.array(m.nu_int_kernel[:nqm])
        d["nu_tau_notmassless"] = np.array([[m.nu_tau_notmassless[q][e] for e in range(ne)] for q in range(nqm)]).reshape(nqm, ne)
        if not ne:
            d["nqmax"] = 0
        d["tau_regions"] = np.array([[r.Low, r.High, r.delta, r.start_index, r.IsLog] for r in m.tau_regions[: m.n_tau_regions]])
        d["tautf"] = np.array(m.tautf[: m.transfer_num_redshifts])
        d["ls"] = ls
        d["kmaxfile"] = kmax.value
        tables.append(ModelTables(d))
    return tables, status


class GalCamb:
    """One GPU context.  Stage methods keep the reference's names."""

    def __init__(self, device=0):
        self.L = load_library()
        self._chk(self.L.galcamb_init_(C.byref(_I(device))))
        self.models = []
        self.nl = 0

    def _chk(self, rc):
        if rc:
            raise GalCambError("galcamb error %d: %s" % (rc, self.L.galcamb_last_error_().decode()))

    def close(self):
        self.L.galcamb_free_()

    # InitSpherBessels, camb/bessels.f90:34
    def InitSpherBessels(self, ls, kmaxfile, AccuracyBoost=1.0):
        ls = np.ascontiguousarray(ls, dtype=np.int32)
        self.nl = len(ls)
        self._chk(self.L.galcamb_set_bessels_(ls.ctypes.data_as(C.POINTER(_I)), C.byref(_I(len(ls))), C.byref(_I(int(kmaxfile))),
                                              C.byref(_D(AccuracyBoost))))

    def bessel_tables(self):
        nx = _I(0)
        self._chk(self.L.galcamb_get_bessels_(C.byref(nx), None, None, None))
        xs = np.zeros(nx.value)
        ajl = np.zeros((self.nl, nx.value))
        ajlpr = np.zeros((self.nl, nx.value))
        self._chk(self.L.galcamb_get_bessels_(C.byref(nx), xs.ctypes.data, ajl.ctypes.data, ajlpr.ctypes.data))
        return xs, ajl, ajlpr

    # CheckLoadedHighLTemplate, camb/modules.f90:1222 ; tmpl[type][l-2], type = TT,EE,TE (+ PP for the lensing stage)
    def set_template(self, tmpl):
        if tmpl is None:
            self._chk(self.L.galcamb_set_template_(None, C.byref(_I(0))))
            self._chk(self.L.galcamb_set_lensing_template_(None, C.byref(_I(0))))
            return
        t = np.ascontiguousarray(tmpl, dtype=np.float64)
        assert t.shape[0] in (3, 4)
        t3 = np.ascontiguousarray(t[:3])
        self._chk(self.L.galcamb_set_template_(t3.ctypes.data, C.byref(_I(t.shape[1] + 1))))
        if t.shape[0] == 4:
            pp = np.ascontiguousarray(t[3])
            self._chk(self.L.galcamb_set_lensing_template_(pp.ctypes.data, C.byref(_I(t.shape[1] + 1))))

    def set_models(self, models):
        self.models = list(models)
        self._chk(self.L.galcamb_batch_begin_(C.byref(_I(len(self.models)))))
        for i, m in enumerate(self.models):
            self._chk(self.L.galcamb_set_model_(C.byref(_I(i)), C.byref(m.cstruct())))

    # hot loop 1, camb/cmbmain.f90:201-206
    def CalcScalarSources(self):
        self._chk(self.L.galcamb_evolve_sources_())

    # InitSourceInterpolation + hot loop 2, camb/cmbmain.f90:244, :263-267
    def SourceToTransfers(self):
        self._chk(self.L.galcamb_los_integrate_())

    # ClTransferToCl, camb/cmbmain.f90:293
    def ClTransferToCl(self):
        self._chk(self.L.galcamb_scalar_cls_())

    # CAMB_TransfersToPowers, camb/camb.f90:87: fast-parameter update on resident transfer functions
    def set_primordial(self, slot, ScalarPowerAmp, an, n_run=0.0, n_runrun=0.0, k_0_scalar=0.05):
        self._chk(self.L.galcamb_set_primordial_(C.byref(_I(slot)), C.byref(_D(ScalarPowerAmp)), C.byref(_D(an)),
                                                 C.byref(_D(n_run)), C.byref(_D(n_runrun)), C.byref(_D(k_0_scalar))))
        d = self.models[slot].d
        d["ScalarPowerAmp"], d["an"], d["n_run"], d["n_runrun"], d["k_0_scalar"] = ScalarPowerAmp, an, n_run, n_runrun, k_0_scalar

    def set_alens(self, slot, ALens=1.0, ALens_Fiducial=0.0):
        """CAMBmain's ALens and lensing's ALens_Fiducial for a resident slot (source/Calculator_CAMB.f90:130-131)."""
        self._chk(self.L.galcamb_set_alens_(C.byref(_I(slot)), C.byref(_D(ALens)), C.byref(_D(ALens_Fiducial))))
        d = self.models[slot].d
        d["ALens"], d["ALens_Fiducial"] = ALens, ALens_Fiducial

    def TransfersToPowers(self, do_lensing=False):
        rc = self.L.galcamb_transfers_to_powers_(C.byref(_I(int(do_lensing))))
        if rc != 8:
            self._chk(rc)

    # lens_Cls, camb/lensing.f90:78 (after ClTransferToCl on DoLensing models)
    def lens_Cls(self):
        rc = self.L.galcamb_lens_cls_()
        if rc != 8:  # 8 = error_norm_lensing in some slots: their rows are NaN, see status()
            self._chk(rc)

    def Cl_lensed(self, slot=0):
        lm = _I(0)
        self._chk(self.L.galcamb_get_lensed_cls_(C.byref(_I(slot)), C.byref(lm), None))
        out = np.zeros((4, lm.value - 1))  # [TT,EE,BB,TE][l-2]
        self._chk(self.L.galcamb_get_lensed_cls_(C.byref(_I(slot)), C.byref(lm), out.ctypes.data))
        return out

    def Src(self, slot=0):
        d = self.models[slot].d
        S = 3 if d["WantLateTime"] else 2
        out = np.zeros((d["nt"], S, d["nk"]))
        self._chk(self.L.galcamb_get_sources_(C.byref(_I(slot)), out.ctypes.data))
        return out

    def put_Src(self, Src, slot=0):
        a = np.ascontiguousarray(Src, dtype=np.float64)
        self._chk(self.L.galcamb_put_sources_(C.byref(_I(slot)), a.ctypes.data))

    def Delta_p_l_k(self, slot=0):
        d = self.models[slot].d
        S = 3 if d["WantLateTime"] else 2
        out = np.zeros((d["nq"], self.nl, S))
        self._chk(self.L.galcamb_get_transfers_(C.byref(_I(slot)), out.ctypes.data))
        return out

    def Cls(self, slot=0):
        d = self.models[slot].d
        nt = 6 if d["WantLateTime"] else 3
        icl = np.zeros((nt, self.nl))
        cl = np.zeros((nt, d["lmax"] - 1))
        self._chk(self.L.galcamb_get_cls_(C.byref(_I(slot)), icl.ctypes.data, cl.ctypes.data))
        return icl, cl

    def put_Cls(self, Cl, slot=0):
        a = np.ascontiguousarray(Cl, dtype=np.float64)
        self._chk(self.L.galcamb_put_cls_(C.byref(_I(slot)), a.ctypes.data))

    # one call for a batch: the public API timed end to end by bench.py
    def batch_cls(self, models, out=None):
        n = len(models)
        arr = (CModel * n)(*[m.cstruct() for m in models])
        self.models = list(models)
        lmax = models[0].d["lmax"]
        if out is None:
            out = np.zeros((n, 3, lmax - 1))
        rc = self.L.galcamb_batch_cls_(arr, C.byref(_I(n)), out.ctypes.data)
        if rc not in (0, 4):  # 4 = error_evolution in some slots: their rows are NaN, see status()
            self._chk(rc)
        return out

    # parameters -> C_l: host pre-stage (CAMBParams_Set ... SetkValuesForInt on the host threads) + stages 1-4 on the device
    def params_to_cls(self, params, nthreads=0, out=None):
        n = len(params)
        arr = (CParams * n)(*params)
        if out is None:
            out = np.zeros((n, 3, params[0].Max_l - 1))
        rc = self.L.galcamb_params_to_cls_(arr, C.byref(_I(n)), C.byref(_I(nthreads)), out.ctypes.data)
        if rc not in (0, 1, 2, 4, 6):  # per-point rejections (reionization, recombination, evolution, background): NaN rows
            self._chk(rc)
        return out

    # pipelined form: prestage_async(next batch) overlaps with prestaged_to_cls(current batch)
    def prestage_async(self, params, nthreads=0):
        n = len(params)
        arr = (CParams * n)(*params)
        self._chk(self.L.galcamb_prestage_async_(arr, C.byref(_I(n)), C.byref(_I(nthreads))))

    def prestage_wait(self):
        return self.L.galcamb_prestage_wait_()  # first per-point rejection id, 0 if none

    def prestaged_to_cls(self, out):
        rc = self.L.galcamb_prestaged_to_cls_(out.ctypes.data)
        if rc not in (0, 4):
            self._chk(rc)
        return out

    def prestaged_upload(self):
        """Tables of the current pre-staged batch -> device (every point must be valid); then the stage methods apply."""
        self._chk(self.L.galcamb_prestaged_upload_())

    # several GPUs from this one process (include/galcamb.h "several GPUs from one process")
    def device_count(self):
        return self.L.galcamb_device_count_()

    def multi_params_to_cls(self, ngpu, params, nthreads=0, out=None):
        """Parameter-batch sharding over ngpu devices of this box: contiguous blocks of the valid points per device."""
        n = len(params)
        arr = (CParams * n)(*params)
        if out is None:
            out = np.zeros((n, 3, params[0].Max_l - 1))
        rc = self.L.galcamb_multi_params_to_cls_(C.byref(_I(ngpu)), arr, C.byref(_I(n)), C.byref(_I(nthreads)), out.ctypes.data)
        if rc not in (0, 1, 2, 4, 6):
            self._chk(rc)
        return out

    def multi_single_cls(self, ngpu, params):
        """One spectrum over ngpu devices (k-modes dealt round-robin, Src and iCl all-reduced with NCCL).  Returns
        (Cl[3, lmax-1], dict of milliseconds)."""
        out = np.zeros((3, params.Max_l - 1))
        ms = np.zeros(4)
        self._chk(self.L.galcamb_multi_single_cls_(C.byref(_I(ngpu)), C.byref(params), out.ctypes.data, ms.ctypes.data))
        return out, dict(zip(("call", "prestage", "evolve", "los"), ms))

    def set_host_threads(self, n):
        self.L.galcamb_set_host_threads_(C.byref(_I(int(n))))

    def set_prestage_device(self, device):
        """Galileon backgrounds of a batch on this device (-1: on the host threads); galcamb_set_prestage_device_."""
        self.L.galcamb_set_prestage_device_(C.byref(_I(int(device))))

    # the same with the lensing post-step: (Cl[n,3,lmax-1], Cl_lensed[n,4,lmax_lensed-1])
    def batch_lensed_cls(self, models):
        n = len(models)
        arr = (CModel * n)(*[m.cstruct() for m in models])
        self.models = list(models)
        lmax = models[0].d["lmax"]
        out = np.zeros((n, 3, lmax - 1))
        buf = np.zeros(n * 4 * (lmax - 1))
        lm = _I(0)
        rc = self.L.galcamb_batch_lensed_cls_(arr, C.byref(_I(n)), out.ctypes.data, C.byref(lm), buf.ctypes.data)
        if rc not in (0, 4, 8):
            self._chk(rc)
        return out, buf[:n * 4 * (lm.value - 1)].reshape(n, 4, lm.value - 1).copy()

    def status(self, slot=0):
        st = _I(0)
        self._chk(self.L.galcamb_get_status_(C.byref(_I(slot)), C.byref(st)))
        return st.value

    def timings(self):
        ms = np.zeros(8)
        self.L.galcamb_get_timings_(ms.ctypes.data)
        return dict(zip("evolve spline_k los cls lensing unused bessel total".split(), ms))

    def transfer(self, slot=0):
        """Matter transfer functions of a slot (galcamb_get_transfer_): dict with q_trans [nq], TransferData [nz][nq][13]
        (= Fortran TransferData(13, num_q_trans, nz), FP64) and sigma8 [nz]."""
        nq, nz = _I(0), _I(0)
        self._chk(self.L.galcamb_get_transfer_(C.byref(_I(slot)), C.byref(nq), C.byref(nz), None, None, None))
        q, T, s8 = np.zeros(nq.value), np.zeros((nz.value, nq.value, 13)), np.zeros(nz.value)
        self._chk(self.L.galcamb_get_transfer_(C.byref(_I(slot)), C.byref(nq), C.byref(nz), q.ctypes.data, T.ctypes.data,
                                                s8.ctypes.data))
        return dict(q_trans=q, TransferData=T, sigma8=s8)

    def host_timings(self):
        ms = np.zeros(8)
        self.L.galcamb_get_host_timings_(ms.ctypes.data)
        return dict(zip("pack h2d_enqueue upload stages d2h collect".split(), ms))

    def evolve_stats(self):
        c = np.zeros(2)
        self.L.galcamb_get_evolve_stats_(c.ctypes.data)
        return dict(zip("rhs_executed outputs_from_k1".split(), c))

    def counters(self):
        c = np.zeros(4)
        self.L.galcamb_get_counters_(c.ctypes.data)
        return dict(zip("n_derivs sum_n n_out n_iter".split(), c))
