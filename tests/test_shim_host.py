"""CPU-only: the SEAL-subset header's host logic (hex strings, Plaintext parsing, parameter streams in none/zlib/zstd,
default tables) through a small C++ program compiled against include/seal/seal.h."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_logic_program():
    from pplp_b200 import build, shim_build
    build.build()
    shim_build.build()
    exe = os.path.join(ROOT, "build", "shim", "host_logic")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "host logic ok" in p.stdout


def test_reference_cmake_project_configures_unmodified():
    """find_package(SEAL 4.1 REQUIRED) in the reference's CMakeLists.txt:29 resolves to cmake/SEALConfig.cmake and all
    five targets link against libpplp_b200.so — no edit to the reference tree."""
    import shutil
    import pytest
    if not os.path.isfile("/root/reference/CMakeLists.txt") or not shutil.which("cmake"):
        pytest.skip("needs /root/reference and cmake (the build container)")
    from pplp_b200 import build, shim_build
    build.build()
    exes = shim_build.build_cmake(force=True)
    assert [os.path.basename(e) for e in exes] == ["pplp", "client", "server", "tc", "ts"]
    cache = open(os.path.join(ROOT, "build", "cmake_ref", "CMakeCache.txt")).read()
    assert "SEAL_DIR:PATH=" + os.path.join(ROOT, "cmake") in cache or "SEAL_DIR:UNINITIALIZED=" + os.path.join(ROOT, "cmake") in cache
    for e in exes:
        assert os.path.exists(e)
        needed = subprocess.run(["readelf", "-d", e], capture_output=True, text=True).stdout
        assert "libpplp_b200.so" in needed, e
