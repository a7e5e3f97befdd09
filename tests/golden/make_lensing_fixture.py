"""Golden output of the oracle's lensing post-step (camb/lensing.f90:107-528) for the config-3 flavour used in the tests:
Galileon point, DoLensing, CosmoMC settings, lmax 2650 (lmax_lensed 2550).  Regression pin; the physical pin is
test_lensed_lcdm_against_reference_theory_cl.  Run from the repo root: python tests/golden/make_lensing_fixture.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_py import Oracle  # noqa: E402

P = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
         c4=-1.03411, cG=0.001188243, DoLensing=1, DoLateRadTruncation=0, lmax=2650)
o = Oracle(**P)
o.run()
o.stage("lensing")
L = o.Cl_lensed()
np.savez_compressed(os.path.join(HERE, "config3_lensing_golden.npz"), params_json=np.frombuffer(json.dumps(P).encode(), dtype=np.uint8),
                    lmax_lensed=o.geti("lmax_lensed"), Cl_lensed_7=L[:, ::7])
print("lmax_lensed", o.geti("lmax_lensed"), L.shape)
