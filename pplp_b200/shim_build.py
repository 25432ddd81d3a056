"""Builds the host-side drop-in artefacts with g++ (no CUDA toolchain needed — they only link libpplp_b200.so):

  build/shim/shim_parity      tests/shim/shim_parity.cc: the reference's call sequence through include/seal/seal.h
  build/dropin/{pplp,client,server,test_client,test_server}
                              the REFERENCE's own drivers, compiled UNMODIFIED from /root/reference/src against
                              include/seal/seal.h (only where /root/reference exists, i.e. the build container; the
                              binaries travel to the GPU box, the sources never enter this repository)
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
REF = "/root/reference"
COMMON = ["-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-L", PKG, "-lpplp_b200", "-lz", "-ldl", "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN/../../pplp_b200"]
# -include cstdint: src/demo.cc includes bloomfilter.h (which uses uint8_t) before any header that declares it; GCC >= 13
# no longer leaks <cstdint> through <sstream>, so the reference needs this one flag on a current toolchain.
DROPIN = {"pplp": "src/demo.cc", "client": "src/client.cc", "server": "src/server.cc", "test_client": "src/test/test_client.cc",
          "test_server": "src/test/test_server.cc"}


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("shim build failed: " + cmd[-1])


def build(force=False):
    hdrs = [os.path.join(ROOT, "include", "seal", "seal.h"), os.path.join(ROOT, "include", "pplp_b200.h")]
    out = os.path.join(ROOT, "build", "shim")
    os.makedirs(out, exist_ok=True)
    built = []
    for name in ("shim_parity", "host_logic"):
        src = os.path.join(ROOT, "tests", "shim", name + ".cc")
        exe = os.path.join(out, name)
        if force or _newer(exe, [src] + hdrs):
            _run(["g++", "-Wall", src] + COMMON + ["-o", exe])
        built.append(exe)
    tools = os.path.join(ROOT, "build", "tools")
    os.makedirs(tools, exist_ok=True)
    src = os.path.join(ROOT, "tools", "batch_server.cc")
    exe = os.path.join(tools, "batch_server")
    if force or _newer(exe, [src] + hdrs):
        _run(["g++", "-Wall", src] + COMMON + ["-o", exe])
    built.append(exe)
    if os.path.isdir(os.path.join(REF, "src")):
        d = os.path.join(ROOT, "build", "dropin")
        os.makedirs(d, exist_ok=True)
        for name, rel in DROPIN.items():
            exe = os.path.join(d, name)
            s = os.path.join(REF, rel)
            if force or _newer(exe, [s] + hdrs):
                _run(["g++", "-w", "-include", "cstdint", "-I", os.path.join(REF, "include"), s] + COMMON + ["-o", exe])
            built.append(exe)
    built += build_cmake(force)
    return built


def build_cmake(force=False):
    """The reference's own CMakeLists.txt, UNMODIFIED, configured with -DSEAL_DIR=<repo>/cmake so that its
    `find_package(SEAL 4.1 REQUIRED)` (CMakeLists.txt:29) resolves to cmake/SEALConfig.cmake; builds pplp, client, server,
    tc, ts into build/cmake_ref/ (only where /root/reference and cmake exist; the binaries travel to the GPU box)."""
    import shutil
    out = os.path.join(ROOT, "build", "cmake_ref")
    names = ["pplp", "client", "server", "tc", "ts"]
    if not os.path.isfile(os.path.join(REF, "CMakeLists.txt")) or not shutil.which("cmake"):
        return [os.path.join(out, x) for x in names if os.path.exists(os.path.join(out, x))]
    deps = [os.path.join(ROOT, "include", "seal", "seal.h"), os.path.join(ROOT, "include", "pplp_b200.h"), os.path.join(ROOT, "cmake", "SEALConfig.cmake"),
            os.path.join(ROOT, "cmake", "SEALConfigVersion.cmake")]
    if force or any(_newer(os.path.join(out, x), deps) for x in names):
        _run(["cmake", "-S", REF, "-B", out, "-DSEAL_DIR=" + os.path.join(ROOT, "cmake"), "-DCMAKE_BUILD_RPATH=" + PKG + ";$ORIGIN/../../pplp_b200"])
        _run(["cmake", "--build", out, "-j", "8"])
    return [os.path.join(out, x) for x in names]


if __name__ == "__main__":
    for b in build(force="--force" in sys.argv):
        print(b)
