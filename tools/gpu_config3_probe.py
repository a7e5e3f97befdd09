"""BASELINE configs[2] (Galileon lensed C_l, AccuracyBoost=2, lmax=5000) on the GPU: stage timings vs batch size, the
oracle (CPU) standing in for the host pre-stage and as the checker of slot 0.  usage: gpu_config3_probe.py [batch ...]"""
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

P = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
         c4=-1.03411, cG=0.001188243, DoLensing=1, AccuracyBoost=2)
o = Oracle(lmax=5000, **P)
t0 = time.time()
st = o.run()
o.stage("lensing")
t_cpu = time.time() - t0
print("oracle: %.2f s per spectrum on %d threads (sources %.2f, los %.2f)" % (t_cpu, os.cpu_count(), st["sources"], st["los"]), flush=True)
t = tables_from_oracle(o, P, lmax=5000)
g = GalCamb(0)
g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())
g.InitSpherBessels(t.ls, t.d["kmaxfile"], 2.0)
rng = np.random.default_rng(20250101)
for B in [int(x) for x in sys.argv[1:]] or [1, 16, 64]:
    models = [t.copy() for _ in range(B)]
    for m in models[1:]:
        m.d["k"] = m.d["k"] * (1 + 0.02 * rng.uniform(-1, 1))  # desynchronise the lanes like distinct cosmologies do
    g.set_models(models)
    t0 = time.time()
    g.CalcScalarSources()
    g.SourceToTransfers()
    g.ClTransferToCl()
    g.lens_Cls()
    wall = time.time() - t0
    tm = g.timings()
    L = g.Cl_lensed(0)
    err = max(float(np.max(np.abs(L[i] / o.Cl_lensed()[i] - 1))) for i in (0, 1, 2))
    print("B=%d wall %.2f s: evolve %.0f ms, spline_k %.1f, los %.1f, cls %.2f, lensing %.2f -> %.2f lensed spectra/s; "
          "max rel err lensed TT/EE/BB slot 0: %.1e" % (B, wall, tm["evolve"], tm["spline_k"], tm["los"], tm["cls"],
                                                         tm["lensing"], B / wall, err), flush=True)
