"""Tool: stage timings vs batch size (replicas of the config-2 table fixture).  usage: gpu_scale_probe.py [--jitter] 32 256 1024
--jitter: every replica but the first gets its source k grid scaled by a random factor within +-2 % (seeded), which
desynchronises the adaptive step sequences of the lanes of a warp the way distinct cosmologies do."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cosmomc_galileon_b200 import GalCamb, ModelTables
H = os.path.join(ROOT, "tests")
t = ModelTables.load(os.path.join(H, "golden", "config2_tables.npz"))
g = GalCamb(0)
tmpl = np.fromfile(os.path.join(H, "golden", "highL_template.f64")).reshape(7999, 4)[:, :3].T.copy()
g.set_template(tmpl)
g.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
gold = np.load(os.path.join(H, "golden", "config2_golden.npz"))
JIT = "--jitter" in sys.argv
rng = np.random.default_rng(20250101)
for B in [int(x) for x in sys.argv[1:] if x != "--jitter"] or [1, 32, 128]:
    models = [t.copy() for _ in range(B)]
    if JIT:
        for m in models[1:]:
            m.d["k"] = m.d["k"] * (1 + 0.02 * rng.uniform(-1, 1))
    t0 = time.time(); g.set_models(models); t_set = time.time() - t0
    t0 = time.time(); g.CalcScalarSources(); g.SourceToTransfers(); g.ClTransferToCl(); wall = time.time() - t0
    tm = g.timings()
    icl, cl = g.Cls(0)
    err = np.max(np.abs(cl / gold["Cl"] - 1))
    print("B=%d set=%.3fs wall=%.3fs evolve=%.1fms splk=%.2f los=%.2f cls=%.2f -> %.2f spectra/s ; Cl err %.2e" % (
        B, t_set, wall, tm["evolve"], tm["spline_k"], tm["los"], tm["cls"], B / wall, err), flush=True)
