"""Timing of the lensing stage alone (stage 5) on the GPU: the same C_l in every slot of a batch.
usage: python tools/gpu_lensing_probe.py [lmax AccuracyBoost batch]..."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cosmomc_galileon_b200 import GalCamb  # noqa: E402
from oracle_py import Oracle  # noqa: E402
from tables_from_oracle import tables_from_oracle  # noqa: E402

GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243, DoLensing=1)
g = GalCamb(0)
g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())
args = [float(a) for a in sys.argv[1:]] or [2650, 1, 1024, 5000, 2, 128]
for lmax, ab, batch in zip(args[0::3], args[1::3], args[2::3]):
    lmax, batch = int(lmax), int(batch)
    oA = Oracle(lmax=lmax, **GAL)
    oA.run()
    P2 = dict(GAL, AccuracyBoost=ab)
    oB = Oracle(lmax=lmax, **P2)
    oB.stage("background", "initvars", "qgrid")
    oB.put("Cl", oA.Cl())
    t0 = time.time()
    oB.stage("lensing")
    t_cpu = time.time() - t0
    t = tables_from_oracle(oB, P2, lmax=lmax)
    g.InitSpherBessels(t.ls, t.d["kmaxfile"], t.d["AccuracyBoost"])
    g.set_models([t] + [t.copy() for _ in range(batch - 1)])
    for i in range(batch):
        g.put_Cls(oA.Cl(), i)
    for rep in range(3):
        g.lens_Cls()
    ms = g.timings()["lensing"]
    L, Lo = g.Cl_lensed(batch - 1), oB.Cl_lensed()
    err = max(float(np.max(np.abs(L[t_] / Lo[t_] - 1))) for t_ in (0, 1, 2))
    print("lmax %d AB %g batch %d: lensing %.2f ms (%.3f ms/model); oracle 1 thread %.1f ms/model; max rel err TT/EE/BB %.2e"
          % (lmax, ab, batch, ms, ms / batch, 1e3 * t_cpu, err), flush=True)
