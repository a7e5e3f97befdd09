"""Tool: stage timings vs batch size (replicas of the config-2 table fixture).  usage: gpu_scale_probe.py 32 256 1024"""
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
for B in [int(x) for x in sys.argv[1:]] or [1, 32, 128]:
    models = [t.copy() for _ in range(B)]
    t0 = time.time(); g.set_models(models); t_set = time.time() - t0
    t0 = time.time(); g.CalcScalarSources(); g.SourceToTransfers(); g.ClTransferToCl(); wall = time.time() - t0
    tm = g.timings()
    icl, cl = g.Cls(B - 1)
    err = np.max(np.abs(cl / gold["Cl"] - 1))
    print("B=%d set=%.3fs wall=%.3fs evolve=%.1fms splk=%.2f los=%.2f cls=%.2f -> %.2f spectra/s ; Cl err %.2e" % (
        B, t_set, wall, tm["evolve"], tm["spline_k"], tm["los"], tm["cls"], B / wall, err), flush=True)
