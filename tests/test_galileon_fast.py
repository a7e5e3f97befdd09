"""The CUDA kernel evaluates the Galileon perturbation terms in a restructured form (galileon_fast.cuh: background
coefficients once per RHS call, terms as linear forms).  That header is plain C++ on the host, so it is checked here,
without a GPU, against the oracle's reference-order evaluation of camb/galileon.cc:746-1081."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restructured_terms_match_reference_order(oracle_lib, tmp_path):
    exe = str(tmp_path / "galfast_check")
    bdir = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests", "galfast_check.cpp"),
                           "-o", exe, "-L" + bdir, "-lorc", "-Wl,-rpath," + bdir, "-fopenmp"])
    out = subprocess.check_output([exe], text=True)
    worst = {l.split()[0]: float(l.split()[-1]) for l in out.strip().splitlines()}
    assert set(worst) == {"dh", "dx", "grhogal", "gpresgal", "Chigal", "qgal", "Pigal", "dphisecond", "pigalprime"}
    for name, e in worst.items():
        assert e < 1e-9, (name, e)  # rounding only (Pigal: ~2e-11 through cancellation)
