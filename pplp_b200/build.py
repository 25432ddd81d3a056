"""Builds pplp_b200/libpplp_b200.so (hand-written sm_100a kernels + the C ABI of include/pplp_b200.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build container; the resulting .so travels to the B200
box with the repository snapshot.  Usage: `python -m pplp_b200.build [--force]`.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libpplp_b200.so")
SOURCES = ["capi.cu", "ntt.cu", "eval.cu", "crypto.cu", "bloom.cu", "behz.cu", "behzf.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
] + os.environ.get("PPLP_NVCC_EXTRA", "").split()   # experiments only (e.g. -DPPLP_NTT_MIN_CTAS=1); part of the staleness digest


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("pplp_b200: nvcc not found; the CUDA extension cannot be built (there is no CPU fallback)")


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".hpp", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "pplp_b200.h"))
    return hs


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(f for f in NVCC_FLAGS if not f.startswith("/")).encode())
    return h.hexdigest()


def _stamp_path(obj):
    return obj + ".sha256"


def _stale(target, deps):
    """Content-hash staleness (mtimes do not survive the snapshot to the GPU box)."""
    if not os.path.exists(target) or not os.path.exists(_stamp_path(target)):
        return True
    return open(_stamp_path(target)).read().strip() != _digest(deps)


def _stamp(target, deps):
    with open(_stamp_path(target), "w") as f:
        f.write(_digest(deps))


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, cmd, obj, [src] + hdrs))

    def run(job):
        name, cmd, obj, deps = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            _stamp(obj, deps)
        return name, r

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {name} ---\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"pplp_b200: nvcc failed on {name}")
    if force or jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("pplp_b200: link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
