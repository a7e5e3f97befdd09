"""GPU bring-up tool (not a pytest): oracle vs CUDA stage by stage, with timings.  Uses the test oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle_py import Oracle, LCDM
from tables_from_oracle import tables_from_oracle
from cosmomc_galileon_b200 import GalCamb

GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243)

def rel(a, b):
    s = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / s

g = GalCamb(0)
tmpl = np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4)[:, :3].T.copy()
g.set_template(tmpl)
for name, P in (("LCDM", LCDM), ("GAL", GAL)):
    o = Oracle(**P)
    t0 = time.time(); o.run(); t_cpu = time.time() - t0
    t = tables_from_oracle(o, P)
    g.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    xs, ajl, ajlpr = g.bessel_tables()
    print(name, "bess x", rel(xs, o.arr("bess_x")), "ajl", rel(ajl.ravel(), o.arr("ajl")), "ajlpr", rel(ajlpr.ravel(), o.arr("ajlpr")))
    g.set_models([t])
    # stage 3 alone on oracle sources
    g.put_Src(o.Src())
    g.SourceToTransfers()
    D = g.Delta_p_l_k(); Do = o.Delta()
    print(name, "LOS on oracle Src: Delta rel", rel(D, Do), g.timings(), g.counters()["n_iter"], o.get("n_iter"))
    g.ClTransferToCl()
    icl, cl = g.Cls()
    print(name, "Cl on oracle Src: iCl", rel(icl, o.iCl()), "Cl", np.max(np.abs(cl / o.Cl() - 1)))
    # full path
    t0 = time.time(); g.CalcScalarSources(); t_ev = time.time() - t0
    S = g.Src(); So = o.Src()
    print(name, "Src rel (max-norm per source):", [rel(S[:, s, :], So[:, s, :]) for s in range(S.shape[1])], "evolve s", t_ev, "cpu total", t_cpu)
    bad = np.unravel_index(np.argmax(np.abs(S - So)), S.shape); print("  worst at (t,s,k)", bad, S[bad], So[bad])
    print("  counters", g.counters(), "oracle n_derivs", o.get("n_derivs"))
    g.SourceToTransfers(); g.ClTransferToCl()
    D = g.Delta_p_l_k()
    icl, cl = g.Cls()
    print(name, "full: Delta rel", rel(D, Do), "Cl maxrel", np.max(np.abs(cl / o.Cl() - 1)), g.timings())
