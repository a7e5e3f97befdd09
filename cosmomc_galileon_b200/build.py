"""Builds libgalcamb_b200.so (hand-written CUDA for sm_100a + the C ABI of include/galcamb.h) in-tree with nvcc.

No JIT cache, no torch extension machinery: the .so sits next to this file so that it travels with the repo.
GALCAMB_FMAD=false builds without FMA contraction: the per-k ODE restates an adaptive integrator whose
accept/reject decisions can then be compared one-to-one with the CPU oracle (built with -ffp-contract=off;
identical RHS-evaluation counts were verified that way).  The default build contracts to FMA (10% faster);
parity stays at the 1e-9 level.  See DESIGN.md "numerics".
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# GALCAMB_VARIANT=<name> builds an experiment variant next to the product library (own object directory,
# libgalcamb_b200_<name>.so); select it at run time with GALCAMB_LIB=<path>
VARIANT = os.environ.get("GALCAMB_VARIANT", "")
OBJ = os.path.join(HERE, "_build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libgalcamb_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
SOURCES = ["evolve.cu", "evolve_batch.cu", "los.cu", "lensing.cu", "capi.cu", "fp64_peak.cu", "galileon_host.cu", "prestage.cu"]
# the two K1 files are also built with -DGC_WITH_TRANSFER (CP%WantTransfer taps); the lane-per-item one also with
# -DGC_EV_NU=1 (records sized for one massive-neutrino eigenstate: +5 % on a 1024-point batch; the eight-lane kernel
# measured 4-10 % slower built that way and keeps the general records)
TRANSFER_SOURCES = ["evolve.cu", "evolve_batch.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FMAD = os.environ.get("GALCAMB_FMAD", "true")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=" + FMAD, "-std=c++17", "-Xcompiler",
         "-fPIC,-fno-math-errno", "-ccbin", "/usr/bin/g++"] + os.environ.get("GALCAMB_NVCC_EXTRA", "").split()


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "galcamb.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "galileon_abi.h"))
    jobs = []
    # (source, object, extra flags): the two K1 files are compiled a second time with the matter-transfer taps
    # galileon_host.cu holds the device copy of the host's background integrator: no FMA contraction there, so that
    # both run the same IEEE arithmetic
    units = [(s, s.replace(".cu", ".o"), ["-fmad=false"] if s == "galileon_host.cu" else []) for s in SOURCES]
    units += [(s, s.replace(".cu", "_tr.o"), ["-DGC_WITH_TRANSFER"]) for s in TRANSFER_SOURCES]
    units += [("evolve_batch.cu", "evolve_batch_nu1.o", ["-DGC_EV_NU=1"])]
    for s, o, extra in units:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, o)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr))
        return r.stderr
    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(OBJ, o) for _, o, _ in units]
    if force or jobs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, "-ccbin", "/usr/bin/g++"] + objs + ["-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
