"""Synthetic MCMC proposal batches for BASELINE configs[1] / configs[3]: parameter points drawn from independent
Gaussians centred on the config-2 point (the in-file test point of camb/galileon.cc:1411-1418), fractional sigma:
omega_b 1 %, omega_c 2 %, h 2 %, tau 10 %, n_s 0.5 %, ln(1e10 A_s) 1 %, c2 / c3 / c4 1 %, cG +-0.02 absolute -- the spread
of a chain near its best fit.  About two thirds of such draws fail the reference's own ghost / Laplace stability
checks (camb/galileon.cc:130-245) and are redrawn by the callers.  Plain dictionaries in CAMB's parameter names: the
GPU arm turns them into galcamb_params (make_params), the CPU reference arm feeds them to its own parameter reader.
"""
import numpy as np

CENTRE = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
              c4=-1.03411, cG=0.001188243, optical_depth=0.09, an=0.96, ScalarPowerAmp=2.1e-9)
SIGMA = (("ombh2", 0.01), ("omch2", 0.02), ("H0", 0.02), ("optical_depth", 0.10), ("an", 0.005), ("c2", 0.01), ("c3", 0.01),
         ("c4", 0.01))


def draw(rng, scale=1.0):
    P = dict(CENTRE)
    for k, s in SIGMA:
        P[k] = CENTRE[k] * (1 + scale * s * rng.standard_normal())
    P["ScalarPowerAmp"] = float(np.exp(np.log(CENTRE["ScalarPowerAmp"] * 1e10) * (1 + scale * 0.01 * rng.standard_normal())) * 1e-10)
    P["cG"] = CENTRE["cG"] + scale * 0.02 * rng.standard_normal()
    return P


def candidates(seed, n, scale=1.0):
    """The first point is the centre itself (the committed golden spectrum belongs to it), then n - 1 seeded draws."""
    rng = np.random.default_rng(seed)
    return [dict(CENTRE)] + [draw(rng, scale) for _ in range(n - 1)]


def valid_points(n, seed, accept, scale=1.0, chunk=None):
    """n points the caller's `accept(list of dict) -> list of bool` keeps, in draw order (deterministic in seed)."""
    rng = np.random.default_rng(seed)
    out, drawn, first = [], 0, True
    while len(out) < n:
        m = chunk or max(16, 3 * (n - len(out)))
        cand = [draw(rng, scale) for _ in range(m)]
        if first:
            cand[0] = dict(CENTRE)
            first = False
        drawn += m
        for P, ok in zip(cand, accept(cand)):
            if ok and len(out) < n:
                out.append(P)
    return out, drawn
