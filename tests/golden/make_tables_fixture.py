"""Generates tests/golden/config2_tables.npz and config2_golden.npz with the CPU oracle.

config2 = BASELINE.json configs[1]: Galileon scalar C_l, lmax = 2500, camb/params.ini accuracy settings.  The
parameter point is the in-file test point of camb/galileon.cc:1411-1418 because the reference's own stability
check (galileon.cc:238) rejects the camb/galileon.ini values (SURVEY.md section 8d fallback rule).
  *_tables.npz : the host-side state after InitVars (what the Fortran host would pass through galcamb_model)
                 + work counters measured by the oracle for this exact input (roofline numerators)
  *_golden.npz : oracle outputs (sampled Src rows, Delta_p_l_k rows, iCl, Cl) the GPU path is compared with
Run in the build container: python tests/golden/make_tables_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle_py import Oracle  # noqa: E402
from tables_from_oracle import tables_from_oracle  # noqa: E402

CONFIG2 = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
               c4=-1.03411, cG=0.001188243)
CONFIG1 = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.00064, H0=70.0)

for name, P in (("config2", CONFIG2), ("config1", CONFIG1)):
    o = Oracle(**P)
    o.run()
    t = tables_from_oracle(o, P)
    S = 2
    nk, nt, nq, nl = t.d["nk"], t.d["nt"], t.d["nq"], len(t.ls)
    n_iter = o.get("n_iter")
    meta = dict(params=P, n_trial=o.get("n_trial"), n_derivs=o.get("n_derivs"), n_out=o.get("n_out"),
                sum_n_trial=o.get("sum_n_trial"), flop_ode=o.get("flop_ode"), n_iter=n_iter,
                flop_los=(12 + 2 * S) * n_iter + 12 * nt * nq + 8 * S * nt * nq,
                bytes_los_operand=n_iter * (32 + 8 * S) + nq * nt * (8 * S + 28),
                bytes_los_compulsory=16 * o.geti("num_xx") * nl + 16 * S * nk * nt + 8 * S * nl * nq)
    t.d["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    if name == "config2":
        t.save(os.path.join(HERE, name + "_tables.npz"))
    Src, D = o.Src(), o.Delta()
    np.savez_compressed(os.path.join(HERE, name + "_golden.npz"), Src_t=Src[::16], Delta_q=D[::32], iCl=o.iCl(), Cl=o.Cl(),
                        k=o.arr("k"), tau=o.arr("tau"), q=o.arr("q"), ls=t.ls,
                        params_json=np.frombuffer(json.dumps(P).encode(), dtype=np.uint8),
                        meta_json=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print(name, meta)
