"""BASELINE configs[3] as the survey words it: a batch of DIFFERENT parameter points
drawn with numpy.random.default_rng(20250101) from independent Gaussians centred on the config-2 point (fractional
sigma: omega_b 1 %, omega_c 2 %, h 2 %, tau 10 %, n_s 0.5 %, ln A_s 1 %, c2/c3/c4 1 %, cG +-0.02 absolute); points the
Galileon stability checks reject are redrawn (count printed).  The host pre-stage (background, RECFAST, thermo
tables, grids) is produced here by the CPU oracle -- this is a measurement tool, not the product path -- and the
GPU path is timed on the distinct tables next to the same number of replicas of one point.
usage: python tools/gpu_hetero_probe.py [n_points] [scale of the sigmas, default 1]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_py import Oracle  # noqa: E402
from tables_from_oracle import tables_from_oracle  # noqa: E402

CENTRE = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
              c4=-1.03411, cG=0.001188243, optical_depth=0.09, an=0.96, ScalarPowerAmp=2.1e-9)


def draw(rng, scale):
    P = dict(CENTRE)
    for k, s in (("ombh2", 0.01), ("omch2", 0.02), ("H0", 0.02), ("optical_depth", 0.10), ("an", 0.005), ("c2", 0.01),
                 ("c3", 0.01), ("c4", 0.01)):
        P[k] = CENTRE[k] * (1 + scale * s * rng.standard_normal())
    P["ScalarPowerAmp"] = float(np.exp(np.log(CENTRE["ScalarPowerAmp"] * 1e10) * (1 + scale * 0.01 * rng.standard_normal())) * 1e-10)
    P["cG"] = CENTRE["cG"] + scale * 0.02 * rng.standard_normal()
    return P


def prestage(P):
    try:
        o = Oracle(**P)
        o.stage("background", "initvars", "qgrid")
        return tables_from_oracle(o, P)
    except Exception:  # noqa: BLE001  rejected by the stability checks (galileon.cc:209-242) or the thermal history
        return None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rng = np.random.default_rng(20250101)
    tables, params, drawn = [], [], 0
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        while len(tables) < n:
            cand = [draw(rng, scale) for _ in range(2 * (n - len(tables)) + 16)]
            drawn += len(cand)
            for P, t in zip(cand, ex.map(prestage, cand)):
                if t is not None and len(tables) < n:
                    tables.append(t)
                    params.append(P)
    t_pre = time.time() - t0
    print("drew %d points for %d accepted (%.1f %% rejected); oracle pre-stage %.1f s on %d threads"
          % (drawn, n, 100.0 * (1 - n / drawn), t_pre, os.cpu_count()), flush=True)
    nk = [t.d["nk"] for t in tables]
    nt = [t.d["nt"] for t in tables]
    print("nk %d..%d  nt %d..%d  nq %d..%d" % (min(nk), max(nk), min(nt), max(nt), min(t.d["nq"] for t in tables),
                                               max(t.d["nq"] for t in tables)), flush=True)
    from cosmomc_galileon_b200 import GalCamb
    g = GalCamb(0)
    g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())
    g.InitSpherBessels(tables[0].ls, 6251, 1.0)  # kmaxfile = int(Max_eta_k)+1 (camb/bessels.f90:60), Max_eta_k = 2.5*lmax
    for name, batch in (("replicas", [tables[0].copy() for _ in range(n)]), ("distinct", tables)):
        for rep in range(2):
            t0 = time.time()
            out = g.batch_cls(batch)
            wall = time.time() - t0
        tm = g.timings()
        bad = sum(1 for i in range(n) if g.status(i) != 0)
        print("%s: evolve %.0f ms, los %.0f ms, total %.0f ms, wall %.2f s -> %.1f spectra/s resident, %.1f e2e; failed slots %d; n_derivs %.4g"
              % (name, tm["evolve"], tm["los"], tm["total"], wall, n / (tm["total"] * 1e-3), n / wall, bad,
                 g.counters()["n_derivs"]), flush=True)
    # parity of a few of the distinct points against their own full oracle run
    worst = 0.0
    for i in (0, n // 2, n - 1):
        o = Oracle(**params[i])
        o.run()
        worst = max(worst, float(np.max(np.abs(out[i] / o.Cl()[:3] - 1)[:2])))
    print("max rel. C_l error (TT, EE) of 3 distinct points vs the oracle: %.2e" % worst, flush=True)


if __name__ == "__main__":
    main()
