"""The C ABI driven from a C99 program (tests/cabi_smoke.c), not through ctypes: header + dlopen + plain pointers, the way the
reference's own native helper is bound (src/algebra/SEQRCS.jl:41-60)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_runs_a_decomposition_through_the_abi(tmp_path):
    exe = str(tmp_path / "cabi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_smoke.c"),
                    "-o", exe, "-ldl", "-lm"], check=True, capture_output=True)
    out = subprocess.run([exe, os.path.join(ROOT, "itensorcpd.jl_b200", "lib", "libitcpd_b200.so")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "CABI_SMOKE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
