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
