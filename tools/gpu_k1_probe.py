"""GPU probe of K1 on the committed config-2 table fixture: parity against the committed golden outputs (no oracle
run needed), work counters and stage timings for a list of batch sizes.
usage: python tools/gpu_k1_probe.py [batch sizes ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmomc_galileon_b200 import GalCamb, ModelTables  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1, 4, 40]
    t = ModelTables.load(os.path.join(GOLD, "config2_tables.npz"))
    gold = np.load(os.path.join(GOLD, "config2_golden.npz"))
    g = GalCamb(0)
    tmpl = np.fromfile(os.path.join(GOLD, "highL_template.f64")).reshape(7999, 4).T.copy()
    g.set_template(tmpl)
    g.InitSpherBessels(t.ls, t.d["kmaxfile"], 1.0)
    for B in sizes:
        t0 = time.time()
        out = g.batch_cls([t.copy() for _ in range(B)])
        wall = time.time() - t0
        err_cl = float(np.max(np.abs(out[0] / gold["Cl"] - 1)))
        S = g.Src(B - 1)[::16]
        err_src = float(np.max(np.abs(S - gold["Src_t"])) / np.max(np.abs(gold["Src_t"])))
        same = all(np.array_equal(out[b], out[0]) for b in range(B))
        tm = g.timings()
        print("B=%d status0=%d Cl err %.3e Src err %.3e identical=%s evolve %.1f ms los %.1f ms wall %.2f s counters %s stats %s"
              % (B, g.status(0), err_cl, err_src, same, tm["evolve"], tm["los"], wall, g.counters(), g.evolve_stats()), flush=True)
    print("golden n_derivs", float(gold["n_derivs"]) if "n_derivs" in gold.files else None)


if __name__ == "__main__":
    main()
