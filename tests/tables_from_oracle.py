"""Builds the host-side table set (cosmomc_galileon_b200.ModelTables) from an oracle model that has run
CAMBParams_Set + InitVars + SetkValuesForSources (+ SetkValuesForInt for q).  TEST INFRASTRUCTURE."""
import numpy as np

from cosmomc_galileon_b200 import ModelTables


def tables_from_oracle(o, params, lmax=2500):
    """o: tests.oracle_py.Oracle after stage('background','initvars') and stage('bessels','los') or at least
    q grid available; params: the dict used to construct it (for flags)."""
    g = o.get
    d = {}
    for n in ("grhom grhog grhor grhob grhoc grhov grhornomass grhok adotrad tau0 taurst taurend tau_maxvis tight_tau "
              "matter_verydom_tau epsw max_etak_scalar").split():
        d[n] = g(n)
    d["AccuracyBoost"] = float(params.get("AccuracyBoost", 1.0))
    d["lAccuracyBoost"] = float(params.get("lAccuracyBoost", 1.0))
    d["use_galileon"] = int(params.get("use_galileon", 0))
    d["DoLateRadTruncation"] = int(params.get("DoLateRadTruncation", 1))
    d["HighAccuracyDefault"] = int(params.get("HighAccuracyDefault", 1))
    d["AccuratePolarization"] = int(params.get("AccuratePolarization", 1))
    d["AccurateReionization"] = int(params.get("AccurateReionization", 1))
    d["WantLateTime"] = int(params.get("DoLensing", 0))
    d["MassiveNuMethod"] = int(params.get("MassiveNuMethod", 1))
    ne = o.geti("Nu_mass_eigenstates")
    d["Nu_mass_eigenstates"] = ne
    d["Num_Nu_massive"] = o.geti("Num_Nu_massive")
    if ne > 0:
        d["nqmax"] = o.geti("nu_nqmax")
        d["grhormass"] = [g("grhormass%d" % (i + 1)) for i in range(ne)]
        d["nu_masses"] = [g("nu_masses%d" % (i + 1)) for i in range(ne)]
        d["nu_q"] = o.arr("nu_q")
        d["nu_int_kernel"] = o.arr("nu_int_kernel")
        d["nu_tau_notmassless"] = np.stack([o.arr("nu_tau_notmassless%d" % (i + 1)) for i in range(ne)], axis=1)  # [iq][eig]
        d["nu_tau_nonrelativistic"] = [g("nu_tau_nonrelativistic%d" % (i + 1)) for i in range(ne)]
        d["nu_tau_massive"] = [g("nu_tau_massive%d" % (i + 1)) for i in range(ne)]
        d["dlnam"] = g("nu_dlnam")
        for n in "r1 p1 dr1 dp1 ddr1 ddp1".split():
            d["nu_" + n] = o.arr("nu_" + n)
    else:
        d["nqmax"] = 0
        d["dlnam"] = 0.0
        for n in ("grhormass", "nu_masses", "nu_q", "nu_int_kernel", "nu_tau_nonrelativistic", "nu_tau_massive"):
            d[n] = []
        d["nu_tau_notmassless"] = np.zeros((0, 0))
        for n in "r1 p1 dr1 dp1 ddr1 ddp1".split():
            d["nu_" + n] = np.zeros(0)
    d["nthermo"] = 20000
    d["tauminn"] = g("tauminn")
    d["dlntau"] = g("dlntau")
    for n in "cs2 dcs2 dotmu ddotmu dddotmu".split():
        d[n] = o.arr(n)
    d["nt"] = o.geti("nt")
    d["tau"] = o.arr("tau")
    d["dtau"] = o.arr("dtau")
    for n in "vis dvis ddvis expmmu opac dopac".split():
        d[n] = o.arr(n)
    d["tau_regions"] = o.arr("regions_tau").reshape(-1, 5)
    if d["use_galileon"]:
        for n in "c2 c3 c4 c5 cG h0 orad".split():
            d[n] = g("gal_" + n)
        n = o.geti("gal_n")
        d["ngal"] = n - 1  # the tabulated nb+1 points; the library appends the extrapolated node itself
        d["gal_a"] = o.arr("gal_a")[: n - 1]
        d["gal_h"] = o.arr("gal_h")[: n - 1]
        d["gal_x"] = o.arr("gal_x")[: n - 1]
    else:
        for n in "c2 c3 c4 c5 cG h0 orad".split():
            d[n] = 0.0
        d["ngal"] = 0
        for n in ("gal_a", "gal_h", "gal_x"):
            d[n] = np.zeros(0)
    d["nk"] = o.geti("nk")
    d["k"] = o.arr("k")
    d["nq"] = o.geti("nq")
    d["q"] = o.arr("q")
    d["dq"] = o.arr("dq")
    d["ScalarPowerAmp"] = float(params.get("ScalarPowerAmp", 2.1e-9))
    d["an"] = float(params.get("an", 0.96))
    d["n_run"] = 0.0
    d["n_runrun"] = 0.0
    d["k_0_scalar"] = 0.05
    d["lmax"] = int(lmax)
    d["ls"] = o.arr("ls").astype(np.int32)
    d["kmaxfile"] = o.geti("kmaxfile")
    return ModelTables(d)
