"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/itcpd_b200.h declares, fails loudly without a GPU (no CPU fallback), and its host-side pieces
(index maps, libc-rand sparse-sign generators) agree with the oracle / the reference's own C."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsparse_sign_ref.so"))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "itcpd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(itcpd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import itcpd

    lib = ctypes.CDLL(itcpd.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/itcpd_b200.h but not exported"
    assert set(itcpd.DECLARED_SYMBOLS) == set(syms), set(itcpd.DECLARED_SYMBOLS) ^ set(syms)
    assert itcpd.load().itcpd_version() >= 100


def test_library_is_sm100a_with_tma_and_dmma():
    """The product .so must contain sm_100a SASS with TMA and FP64 tensor-core instructions."""
    import itcpd

    try:
        elf = subprocess.run(["cuobjdump", "-lelf", itcpd.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", itcpd.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "DMMA.8x8x4" in sass and "UTMALDG" in sass and "UBLKCP" in sass


def test_no_cpu_fallback_without_gpu():
    import itcpd

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(itcpd.ItcpdError) as e:
        itcpd.Engine(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(itcpd.ItcpdError):
        itcpd.decompose(np.zeros((3, 4, 5)), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "itensorcpd.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or "import oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt, f


def test_index_maps_match_oracle_and_reference_literals():
    import itcpd
    from oracle import sampled

    cols = [1, 4, 7, 12, 29, 30, 8, 21, 17, 42, 62, 86, 72]  # test/pivot_mapping.jl:17
    a = itcpd.column_to_multi_coords(cols, (5, 6, 3))
    assert np.array_equal(a, sampled.column_to_multi_coords(cols, (5, 6, 3)))
    assert np.array_equal(itcpd.multi_coords_to_column((5, 6, 3), a), np.array(cols))
    rng = np.random.default_rng(0)
    cols = rng.integers(1, 7 * 11 * 13 * 2 + 1, size=500)
    a = itcpd.column_to_multi_coords(cols, (7, 11, 13, 2))
    assert np.array_equal(a, sampled.column_to_multi_coords(cols, (7, 11, 13, 2)))
    assert np.array_equal(itcpd.multi_coords_to_column((7, 11, 13, 2), a), cols)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("inj", [False, True])
@pytest.mark.parametrize("l,n,s", [(1200, 10000, 8), (50, 31, 2), (40, 62, 1), (7, 5, 9)])
def test_library_sparse_sign_bit_exact_vs_reference_c(inj, l, n, s):
    """itcpd_sparse_sign / itcpd_sparsestack keep the reference's C ABI and libc rand() stream."""
    import itcpd
    from oracle import sampled

    libc = ctypes.CDLL(None)
    rows, vals, cs = itcpd.sparse_sign_matrix(l, n, s, injective=inj, seed=99)
    after = [libc.rand() for _ in range(40)]
    rv, rr, rc = sampled.sparse_sign_call(l, n, s, inj, "ref", seed=99)
    after_ref = [libc.rand() for _ in range(40)]
    assert np.array_equal(rows, rr) and np.array_equal(cs, rc) and np.array_equal(vals, rv, equal_nan=True)
    assert after == after_ref      # the global stream continues exactly where the reference's C would have left it


def test_lock_free_rand_stream_is_in_use_and_equals_plain_rand(tmp_path):
    """The generators borrow glibc's state array instead of calling rand() per draw (sparse_sign.cu: GlibcStream).  On this image's
    glibc the self-test must accept it (else the set-up timings in DESIGN.md are not what runs), and a process forced onto plain
    rand() (ITCPD_PLAIN_RAND) must produce the same arrays and leave the same stream behind, across consecutive calls."""
    import itcpd

    assert itcpd.load().itcpd_sparse_sign_fast_stream() == 1
    prog = (
        "import sys, ctypes, hashlib, numpy as np; sys.path.insert(0, %r); import itcpd\n"
        "libc = ctypes.CDLL(None); libc.srand(2026); h = hashlib.sha256()\n"
        "for inj, l, n, s in [(False, 300, 5000, 4), (True, 300, 5000, 4), (False, 9, 700, 9), (True, 64, 3, 5), (False, 31, 31 * 31, 1)]:\n"
        "    rows, vals, cs = itcpd.sparse_sign_matrix(l, n, s, injective=inj)\n"
        "    h.update(rows.tobytes()); h.update(np.nan_to_num(vals).tobytes()); h.update(str(libc.rand()).encode())\n"
        "print(itcpd.load().itcpd_sparse_sign_fast_stream(), h.hexdigest())\n" % ROOT)
    outs = []
    for env_extra in ({}, {"ITCPD_PLAIN_RAND": "1"}):
        env = dict(os.environ, **env_extra)
        out = subprocess.run(["python", "-c", prog], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout.split())
    assert outs[0][0] == "1" and outs[1][0] == "0" and outs[0][1] == outs[1][1]


def test_host_fitcheck_matches_oracle_state_machine():
    import itcpd
    from oracle import cpals

    rng = np.random.default_rng(1)
    a, b = itcpd.FitCheck(1e-3, 7, 12.5), cpals.FitCheck(1e-3, 7, 12.5)
    for _ in range(3):  # re-used checks (README.md:100-103 behaviour: lastfit resets to 0)
        for _ in range(9):
            inner, sq = float(rng.uniform(40, 60)), float(rng.uniform(20, 30))
            ra = a.update(inner, sq, 4)
            b.iter += 1
            rb = b.update(inner, sq, 4)
            assert ra == rb and a.iter == b.iter and a.counter == b.counter and a.lastfit == b.lastfit
            assert a.final_fit == b.final_fit and a.total_iter == b.total_iter


def test_host_random_cpd_matches_oracle_shape_of_computation():
    import itcpd
    from oracle import cpals

    a = itcpd.random_CPD((5, 6, 7), 4, np.random.default_rng(3))
    b = cpals.random_CPD((5, 6, 7), 4, np.random.default_rng(3))
    for x, y in zip(a.factors, b.factors):
        assert np.array_equal(x, y)
    assert np.array_equal(a.lam, b.lam)
    assert a.rank == 4 and a.dims == (5, 6, 7) and a[()] is a.lam and a[1] is a.factors[1]
    c = itcpd.increase_cpd_rank(a, 6, np.random.default_rng(4))
    assert c.rank == 6 and np.array_equal(c.factors[0][:, :4], a.factors[0])


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run(["python", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "S", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json

    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_build_recipes_cover_every_translation_unit():
    """csrc/build.sh and the Julia deps/build_b200.jl must compile the same, complete set of .cu files (a stale list links with
    undefined symbols only on the maintainer's machine)."""
    csrc = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc")
    units = sorted(f[:-3] for f in os.listdir(csrc) if f.endswith(".cu"))
    sh = open(os.path.join(csrc, "build.sh")).read()
    listed = re.search(r"for f in ([a-z0-9_ ]+); do", sh).group(1).split()
    assert sorted(listed) == units
    jl = open(os.path.join(ROOT, "itensorcpd.jl_b200", "julia", "deps", "b200_paths.jl")).read()   # the one build recipe both Julia files use
    assert 'readdir(csrc)' in jl and 'endswith(".cu")' in jl
    ext = open(os.path.join(ROOT, "itensorcpd.jl_b200", "julia", "ext", "ITCPDB200Ext", "ITCPDB200Ext.jl")).read()
    # every entry point the extension ccalls is declared in the header
    called = set(re.findall(r"ccall\(\(:(itcpd_[a-z0-9_]+), libitcpd\)", ext))
    assert len(called) >= 12 and called <= set(header_symbols()), called - set(header_symbols())


def test_c99_program_compiles_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/cabi_smoke.c is the ABI exercised from C (dlopen + plain pointers).  Here (no GPU) it must compile as strict C99
    against include/itcpd_b200.h, resolve every symbol it uses, and stop at itcpd_create with ITCPD_ERR_NO_DEVICE -- there is
    no CPU fallback to fall into.  The GPU suite runs the same program to completion (tests/test_gpu_cabi_c.py)."""
    exe = str(tmp_path / "cabi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_smoke.c"),
                    "-o", exe, "-ldl", "-lm"], check=True, capture_output=True)
    out = subprocess.run([exe, os.path.join(ROOT, "itensorcpd.jl_b200", "lib", "libitcpd_b200.so")], capture_output=True, text=True, timeout=120)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert out.returncode == 0 and "CABI_SMOKE_OK" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 77 and "no CPU fallback" in out.stderr, (out.returncode, out.stdout, out.stderr)
