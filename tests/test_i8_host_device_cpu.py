"""The REAL host code of the INT8 contraction driving the VERBATIM device code, without a GPU.

tests/test_i8_kernel_simulation_cpu.py runs the gemm_i8.cu kernels on a functional tcgen05 / TMEM / TMA model, but from its own
harness; tests/test_dry_run_cpu.py runs the library's host code, but executes no kernel.  Here the two are joined: the dry-run
library (product objects, `-cudart shared`, behind tests/fake_cudart.cpp) gets a launch hook, a plug-in built from the same
functional model, which EXECUTES every kernel of the INT8 path -- exponent pre-passes, Khatri-Rao digit packing, tensor digit
packing, both GEMM variants, the split-K fix-up -- from the argument block that `launch_partial_gemm_i8` really passed
(cache handling, rank-block strides, partial-tile buffers, TMA descriptor).  The hardware-validated FP64 kernels around it
(second-level contraction, Khatri-Rao expansion, padded upload) are stood in for by plain loops restating their documented
formulas.  `itcpd_mttkrp` through the C-ABI must then match the FP64 MTTKRP to the product's 1e-12 bar."""
import json
import os
import shutil
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_i8_kernel_simulation_cpu import model_source  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRY = os.path.join(ROOT, "oracle", "_build", "dry")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None,
                                reason="needs the CUDA toolkit to link the dry-run copy of the library; no GPU")

HOOK = r"""
// ---------------------------------------------------------------------------------------------------------------------------
// launch hook: kernel name + argument block of the real launch -> emulated execution
// ---------------------------------------------------------------------------------------------------------------------------
#include <string>
#define ARG(T, i) (*(T *)args[i])
struct KrpArgsRef { const double *fac[ITCPD_MAX_ORDER]; int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER]; int nf; int64_t kext; int R; };

template <class F>
static void run_grid(unsigned gx, unsigned gy, unsigned threads, F body) {
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch((int)threads, 0, body, bx, by, gx, gy);
}
template <class F>
static void run_gemm_grid(unsigned gx, unsigned threads, F body) {
    for (unsigned bx = 0; bx < gx; ++bx) {
        g_bars.clear();
        memset(TMEM, 0x5a, sizeof(TMEM));      // stale accumulators must not leak into results
        emu_launch((int)threads, 0, body, bx, 0, gx, 1);
    }
}

extern "C" int emu_launch_hook(const char *name, void **args, unsigned gx, unsigned gy, unsigned gz, unsigned bx_, unsigned by_, unsigned bz_, size_t) {
    const std::string n(name);
    const unsigned threads = bx_ * by_ * bz_;
    auto has = [&](const char *s) { return n.find(s) != std::string::npos; };
    (void)gz;
    // ---- the INT8 path: verbatim kernels ----
    if (has("i8_fill_int_kernel")) { run_grid(gx, gy, threads, [&] { i8_fill_int_kernel(ARG(int *, 0), ARG(int64_t, 1), ARG(int, 2)); }); return 1; }
    if (has("i8_row_exponent_strided_kernel")) {
        run_grid(gx, gy, threads, [&] { i8_row_exponent_strided_kernel(ARG(const double *, 0), ARG(int64_t, 1), ARG(int64_t, 2), ARG(int64_t, 3), ARG(int *, 4)); });
        return 1;
    }
    if (has("i8_row_exponent_contig_kernel")) {
        run_grid(gx, gy, threads, [&] { i8_row_exponent_contig_kernel(ARG(const double *, 0), ARG(int64_t, 1), ARG(int64_t, 2), ARG(int64_t, 3), ARG(int *, 4)); });
        return 1;
    }
    if (has("i8_krp_exponent_kernel")) { run_grid(gx, gy, threads, [&] { i8_krp_exponent_kernel(ARG(I8Krp, 0), ARG(int *, 1)); }); return 1; }
    if (has("i8_krp_pack_kernel")) { run_grid(gx, gy, threads, [&] { i8_krp_pack_kernel(ARG(I8Krp, 0), ARG(const int *, 1), ARG(int64_t, 2), ARG(uint8_t *, 3)); }); return 1; }
    if (has("i8_pack_tensor_kernelILi0E") || has("i8_pack_tensor_kernelILi1E")) {
        const bool k0 = has("ILi0E");
        run_grid(gx, gy, threads, [&] {
            if (k0) i8_pack_tensor_kernel<0>(ARG(const double *, 0), ARG(int64_t, 1), ARG(int64_t, 2), ARG(int64_t, 3), ARG(int64_t, 4), ARG(const int *, 5), ARG(int64_t, 6), ARG(uint8_t *, 7));
            else i8_pack_tensor_kernel<1>(ARG(const double *, 0), ARG(int64_t, 1), ARG(int64_t, 2), ARG(int64_t, 3), ARG(int64_t, 4), ARG(const int *, 5), ARG(int64_t, 6), ARG(uint8_t *, 7));
        });
        return 1;
    }
    if (has("partial_gemm_i8_kernelILi0E") || has("partial_gemm_i8_kernelILi1E")) {
        const bool k0 = has("ILi0E");
        run_gemm_grid(gx, threads, [&] {
            if (k0) partial_gemm_i8_kernel<0>(ARG(CUtensorMap, 0), ARG(const uint8_t *, 1), ARG(const int *, 2), ARG(const int *, 3), ARG(double *, 4), ARG(int64_t, 5), ARG(int, 6), ARG(I8Sched, 7), ARG(double *, 8));
            else partial_gemm_i8_kernel<1>(ARG(CUtensorMap, 0), ARG(const uint8_t *, 1), ARG(const int *, 2), ARG(const int *, 3), ARG(double *, 4), ARG(int64_t, 5), ARG(int, 6), ARG(I8Sched, 7), ARG(double *, 8));
        });
        return 1;
    }
    if (has("partial_gemm_i8p_kernelILi0E") || has("partial_gemm_i8p_kernelILi1E")) {
        const bool k0 = has("ILi0E");
        run_gemm_grid(gx, threads, [&] {
            if (k0) partial_gemm_i8p_kernel<0>(ARG(const uint8_t *, 0), ARG(const uint8_t *, 1), ARG(const int *, 2), ARG(const int *, 3), ARG(double *, 4), ARG(int64_t, 5), ARG(int, 6), ARG(I8Sched, 7), ARG(double *, 8));
            else partial_gemm_i8p_kernel<1>(ARG(const uint8_t *, 0), ARG(const uint8_t *, 1), ARG(const int *, 2), ARG(const int *, 3), ARG(double *, 4), ARG(int64_t, 5), ARG(int, 6), ARG(I8Sched, 7), ARG(double *, 8));
        });
        return 1;
    }
    if (has("i8_splitk_fixup_kernel")) {
        run_grid(gx, gy, threads, [&] { i8_splitk_fixup_kernel(ARG(const double *, 0), ARG(int, 1), ARG(int, 2), ARG(int64_t, 3), ARG(int64_t, 4), ARG(double *, 5)); });
        return 1;
    }
    // ---- hardware-validated FP64 kernels around it: plain loops restating their documented formulas (kernels.cu) ----
    if (has("krp_expand_kernel")) {            // W[k, r] = prod_f A_f[i_f(k), r], zero on padded rows
        const KrpArgsRef a = ARG(KrpArgsRef, 0);
        double *W = ARG(double *, 1);
        for (int r = 0; r < a.R; ++r)
            for (int64_t k = 0; k < a.kext; ++k) {
                double v = 1.0;
                int64_t rem = k;
                for (int f = 0; f < a.nf; ++f) {
                    const int64_t i = rem % a.ext[f];
                    rem /= a.ext[f];
                    v = (i < a.dim[f]) ? v * a.fac[f][i + a.dim[f] * (int64_t)r] : 0.0;
                }
                W[k + a.kext * r] = v;
            }
        return 1;
    }
    if (has("partial_first_kernel")) {         // out[i, r] = sum_b wb[b, r] P[i + I (b + B r)]
        const double *P = ARG(const double *, 0), *wb = ARG(const double *, 1);
        const int64_t I = ARG(int64_t, 2), Ilog = ARG(int64_t, 3), B = ARG(int64_t, 4);
        double *out = ARG(double *, 5);
        for (unsigned r = 0; r < gy; ++r)
            for (int64_t i = 0; i < Ilog; ++i) {
                long double s = 0;
                for (int64_t b = 0; b < B; ++b) s += (long double)P[i + I * (b + B * (int64_t)r)] * (wb ? wb[b + B * (int64_t)r] : 1.0);
                out[i + Ilog * (int64_t)r] = (double)s;
            }
        return 1;
    }
    if (has("partial_general_kernel")) {       // out[i, r] = sum_b wb[b, r] sum_f wf[f, r] P[f + F (i + I (b + B r))]
        const double *P = ARG(const double *, 0), *wf = ARG(const double *, 1), *wb = ARG(const double *, 2);
        const int64_t F = ARG(int64_t, 3), I = ARG(int64_t, 4), B = ARG(int64_t, 5);
        double *out = ARG(double *, 6);
        for (unsigned r = 0; r < gy; ++r)
            for (int64_t i = 0; i < I; ++i) {
                long double s = 0;
                for (int64_t b = 0; b < B; ++b) {
                    long double in = 0;
                    for (int64_t f = 0; f < F; ++f) in += (long double)P[f + F * (i + I * (b + B * (int64_t)r))] * wf[f + F * (int64_t)r];
                    s += in * (wb ? wb[b + B * (int64_t)r] : 1.0);
                }
                out[i + I * (int64_t)r] = (double)s;
            }
        return 1;
    }
    if (has("pad_in_kernel")) {                // dst[i0 + ld0 rest] = i0 < d0 ? src[i0 + d0 rest] : 0
        const double *src = ARG(const double *, 0);
        double *dst = ARG(double *, 1);
        const int64_t d0 = ARG(int64_t, 2), ld0 = ARG(int64_t, 3), nstore = ARG(int64_t, 4);
        for (int64_t j = 0; j < nstore; ++j) { const int64_t i0 = j % ld0, rest = j / ld0; dst[j] = i0 < d0 ? src[i0 + d0 * rest] : 0.0; }
        return 1;
    }
    return 0;   // not on this path: the launch stays a dry run
}
"""

DRIVER = r"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = sys.argv[1]; DRY = os.path.join(ROOT, "oracle", "_build", "dry")
sys.path.insert(0, ROOT)
import itcpd
itcpd.package._lib.LIB_PATH = os.path.join(DRY, "libitcpd_dry.so")
fake = C.CDLL(os.path.join(DRY, "libcudart.so.12"))
plug = C.CDLL(os.path.join(DRY, "i8_hook.so"))
fake.fakecuda_set_real_copy_limit.argtypes = [C.c_ulonglong]
fake.fakecuda_set_real_copy_limit(1 << 30)
fake.fakecuda_set_launch_hook.argtypes = [C.c_void_p]
fake.fakecuda_set_launch_hook(C.cast(plug.emu_launch_hook, C.c_void_p))
fake.fakecuda_launches.restype = C.c_long

def mttkrp_ref(T, fac, n):
    N = T.ndim
    letters = "abcdefgh"[:N]
    ops, sub = [T.astype(np.longdouble)], [letters]
    for m in range(N):
        if m != n:
            ops.append(fac[m].astype(np.longdouble)); sub.append(letters[m] + "r")
    return np.einsum(",".join(sub) + "->" + letters[n] + "r", *ops).astype(np.float64)

out = []
cases = json.loads(sys.argv[2])
with itcpd.Engine(0) as eng:
    for case in cases:
        dims, R, variant, sa, sb = tuple(case["dims"]), case["R"], case["variant"], case["split_a"], case["split_b"]
        rng = np.random.default_rng(case["seed"])
        T = np.asfortranarray(rng.standard_normal(dims) * np.exp2(rng.integers(-6, 7, size=(dims[0],) + (1,) * (len(dims) - 1))))
        fac = [np.asfortranarray(rng.standard_normal((d, R))) for d in dims]
        fake.fakecuda_clear()
        eng.set_option("gemm_i8", variant); eng.set_option("split_a", sa); eng.set_option("split_b", sb)
        eng.set_tensor(T); eng.set_cpd(fac, np.ones(R))
        errs = []
        for rep in range(2):            # the second round re-uses the cached row exponents / digit planes with NEW factors
            for n in range(len(dims)):
                M = eng.mttkrp(n)
                ref = mttkrp_ref(T, fac, n)
                errs.append(float(np.linalg.norm(M - ref) / np.linalg.norm(ref)))
            fac = [np.asfortranarray(rng.standard_normal((d, R))) for d in dims]
            eng.set_cpd(fac, np.ones(R))
        nv = fake.fakecuda_violation_count()
        out.append({"case": case, "max_err": max(errs), "violations": nv,
                    "i8_gemm": int(fake.fakecuda_launches(b"<executed> _ZN5itcpd22partial_gemm_i8_kernel")) + int(fake.fakecuda_launches(b"<executed> _ZN5itcpd23partial_gemm_i8p_kernel")),
                    "dmma": int(fake.fakecuda_launches(b"partial_gemm_kernel")), "fixup": int(fake.fakecuda_launches(b"<executed> _ZN5itcpd22i8_splitk_fixup"))})
print("I8HD_JSON " + json.dumps(out))
"""

CASES = [
    # dims, R, forced splits: every case goes through launch_partial_gemm_i8 for both passes
    {"dims": [128, 12, 10], "R": 20, "split_a": 2, "split_b": 1},      # (2,1) tree: P_A, P_B + second-level contractions
    {"dims": [64, 16, 40], "R": 70, "split_a": 1, "split_b": 1},       # (1,1): pass A is M_0 itself, 2 rank blocks, split-K (1 row tile x 20 k-tiles)
    {"dims": [37, 9, 11], "R": 16, "split_a": 2, "split_b": 2},        # odd leading mode (padded storage), ragged tiles
    {"dims": [16, 12, 6, 5], "R": 130, "split_a": 2, "split_b": 2},    # order 4, 3 rank blocks
    {"dims": [40, 50], "R": 8, "split_a": 1, "split_b": 1},            # order 2
]


@pytest.fixture(scope="module")
def results():
    # the dry-run library, kernel table and fake runtime (same recipe as tests/test_dry_run_cpu.py)
    from test_dry_run_cpu import CUDA_INC, OBJ
    subprocess.run(["bash", os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "build.sh")], check=True, capture_output=True)
    os.makedirs(DRY, exist_ok=True)
    objs = sorted(os.path.join(OBJ, f) for f in os.listdir(OBJ) if f.endswith(".o"))
    subprocess.run(["nvcc", "-shared", "-o", os.path.join(DRY, "libitcpd_dry.so"), *objs, "-cudart", "shared", "-ldl", "-lpthread", "-lrt"], check=True, capture_output=True)
    vermap = os.path.join(DRY, "ver.map")
    open(vermap, "w").write("libcudart.so.12 { global: *; };\n")
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "fake_cudart.cpp"), "-I", CUDA_INC,
                    f"-Wl,--version-script={vermap}", "-Wl,-soname,libcudart.so.12", "-o", os.path.join(DRY, "libcudart.so.12")], check=True, capture_output=True)
    cpp, so = os.path.join(DRY, "i8_hook.cpp"), os.path.join(DRY, "i8_hook.so")
    open(cpp, "w").write(model_source() + HOOK)
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "tests"), "-o", so, cpp, "-lpthread"],
                   check=True, capture_output=True)
    cases = [dict(c, variant=v, seed=11 + i) for i, c in enumerate(CASES) for v in (1, 2)]
    env = dict(os.environ, LD_LIBRARY_PATH=DRY)
    env.pop("FAKECUDA_KERNEL_TABLE", None)
    for k in ("ITCPD_GEMM_I8", "ITCPD_EARLY_B", "ITCPD_CHOL", "ITCPD_NO_GRAPH"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, "-c", DRIVER, ROOT, json.dumps(cases)], env=env, capture_output=True, text=True, timeout=1500)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("I8HD_JSON ")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-1500:], p.stderr[-3000:])
    return json.loads(lines[-1][len("I8HD_JSON "):])


def test_real_host_code_drives_the_emulated_int8_kernels(results):
    assert len(results) == 2 * len(CASES)
    for r in results:
        assert r["violations"] == 0, r
        assert r["i8_gemm"] >= 2 and r["dmma"] == 0, r            # both passes went through the INT8 kernels, never the DMMA fallback
        assert r["max_err"] < 1e-12, r
    assert any(r["fixup"] > 0 for r in results)                    # the split-K schedule was exercised end to end
