"""The hand-encoded UMMA descriptors of csrc/gemm_i8.cu against CuTe's own constructors (the CUTLASS headers vendored with
flashinfer): `make_umma_desc` must accept the kernel's shared-memory digit-plane layouts as canonical (its static_asserts)
and produce the same 64-bit matrix descriptors, and `make_instr_desc` the same 32-bit instruction descriptors.  Host-only
compile with nvcc; skipped when the headers are not installed."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "gemm_i8.cu")
BUILD = os.path.join(ROOT, "oracle", "_build")

PROGRAM = r"""
#include <cstdio>
#include <cstdint>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;
%(funcs)s
template <class L> static uint64_t cute_desc_k(L l) { return (uint64_t)UMMA::make_umma_desc<UMMA::Major::K>(make_tensor(make_smem_ptr((int8_t *)nullptr), l)); }
template <class L> static uint64_t cute_desc_mn(L l) { return (uint64_t)UMMA::make_umma_desc<UMMA::Major::MN>(make_tensor(make_smem_ptr((int8_t *)nullptr), l)); }
int main() {
    int bad = 0;
    // KIND 0 A plane, 128 x 32 int8, byte(m, k) = (m%%8)*16 + (m/8)*256 + (k/16)*128 + k%%16
    auto lk = make_layout(make_shape(make_shape(Int<8>{}, Int<16>{}), make_shape(Int<16>{}, Int<2>{})),
                          make_stride(make_stride(Int<16>{}, Int<256>{}), make_stride(Int<1>{}, Int<128>{})));
    bad += cute_desc_k(lk) != i8_smem_desc(0, 128, 256);
    // stacked B planes, 448 x 32 int8, same layout
    auto lb = make_layout(make_shape(make_shape(Int<8>{}, Int<56>{}), make_shape(Int<16>{}, Int<2>{})),
                          make_stride(make_stride(Int<16>{}, Int<256>{}), make_stride(Int<1>{}, Int<128>{})));
    bad += cute_desc_k(lb) != i8_smem_desc(0, 128, 256);
    // KIND 1 A plane, MN-major: byte(n, k) = (k%%8)*16 + (k/8)*128 + (n/16)*512 + n%%16
    auto lm = make_layout(make_shape(make_shape(Int<16>{}, Int<8>{}), make_shape(Int<8>{}, Int<4>{})),
                          make_stride(make_stride(Int<1>{}, Int<512>{}), make_stride(Int<16>{}, Int<128>{})));
    bad += cute_desc_mn(lm) != i8_smem_desc(0, 128, 512);
    printf("desc K %%016llx  B %%016llx  MN %%016llx\n", (unsigned long long)cute_desc_k(lk), (unsigned long long)cute_desc_k(lb), (unsigned long long)cute_desc_mn(lm));
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 256, UMMA::Major::K, UMMA::Major::K>() != i8_idesc(256, 0);
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 192, UMMA::Major::K, UMMA::Major::K>() != i8_idesc(192, 0);
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 128, UMMA::Major::K, UMMA::Major::K>() != i8_idesc(128, 0);
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 64, UMMA::Major::K, UMMA::Major::K>() != i8_idesc(64, 0);
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 256, UMMA::Major::MN, UMMA::Major::K>() != i8_idesc(256, 1);
    bad += (uint32_t)UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 64, UMMA::Major::MN, UMMA::Major::K>() != i8_idesc(64, 1);
    printf(bad ? "DESC_MISMATCH %%d\n" : "DESC_OK\n", bad);
    return bad;
}
"""


def test_hand_encoded_umma_descriptors_match_cute():
    import importlib.util

    spec = importlib.util.find_spec("flashinfer")   # located, not imported
    if spec is None or not spec.origin:
        pytest.skip("flashinfer (vendored CUTLASS headers) not installed")
    inc = os.path.join(os.path.dirname(spec.origin), "data", "cutlass", "include")
    if not os.path.exists(os.path.join(inc, "cute", "atom", "mma_traits_sm100.hpp")):
        pytest.skip("CuTe sm100 headers not found")
    text = open(SRC).read()
    f1 = re.search(r"__host__ __device__ __forceinline__ uint64_t i8_smem_desc\(.*?\n\}\n", text, flags=re.S).group(0)
    f2 = re.search(r"__host__ __device__ constexpr uint32_t i8_idesc\(.*?\n\}\n", text, flags=re.S).group(0)
    os.makedirs(BUILD, exist_ok=True)
    cu, exe = os.path.join(BUILD, "i8_desc_check.cu"), os.path.join(BUILD, "i8_desc_check")
    open(cu, "w").write(PROGRAM % {"funcs": f1 + f2})
    subprocess.run(["nvcc", "-std=c++17", "-I", inc, "-arch=sm_100a", "--expt-relaxed-constexpr", "-o", exe, cu], check=True, capture_output=True,
                   timeout=900)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "DESC_OK" in out.stdout, out.stdout + out.stderr[-2000:]
