"""The stable, multi-threaded counting sort that turns the sparse-sign embedding into CSR-by-sketch-row (api.cu: sketch_csr_fill) is
pure host code: compile it as it stands (text taken from api.cu) and compare it with a serial stable sort for several thread counts.
The order inside a row is the summation order of the sketch kernel, so it must be exactly the serial one."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MAIN = r"""
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
enum { ITCPD_OK = 0, ITCPD_ERR_ARG = 2 };
%s
int main() {
    unsigned st = 12345;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return st >> 8; };
    const int threads[] = {1, 2, 3, 5, 8, 64};
    for (int trial = 0; trial < 40; ++trial) {
        const int l = 1 + rnd() %% 200, s_eff = 1 + rnd() %% 7;
        const int64_t ncols = 1 + rnd() %% 3000, nnz = ncols * s_eff;
        std::vector<int> rows((size_t)nnz);
        std::vector<double> vals((size_t)nnz);
        for (auto &r : rows) r = (int)(rnd() %% l);
        for (auto &v : vals) v = (double)rnd();
        // serial reference: stable sort of the non-zero numbers by row
        std::vector<int64_t> order((size_t)nnz);
        for (int64_t q = 0; q < nnz; ++q) order[q] = q;
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rows[a] < rows[b]; });
        for (int nt : threads) {
            char buf[16];
            snprintf(buf, sizeof buf, "%%d", nt);
            setenv("ITCPD_SKETCH_THREADS", buf, 1);
            SketchCsr k;
            k.col_own.assign((size_t)nnz, -1);
            k.val_own.assign((size_t)nnz, -1.0);
            k.col = k.col_own.data();
            k.val = k.val_own.data();
            if (sketch_csr_fill(l, s_eff, ncols, rows.data(), vals.data(), k) != ITCPD_OK) return 2;
            if (k.row_ptr[0] != 0 || k.row_ptr[l] != nnz) return 3;
            for (int64_t pos = 0; pos < nnz; ++pos) {
                const int64_t q = order[pos];
                if (k.col[pos] != q / s_eff || k.val[pos] != vals[q]) { printf("mismatch trial %%d nt %%d pos %%lld\n", trial, nt, (long long)pos); return 4; }
            }
            for (int r = 0; r < l; ++r)
                for (int64_t e = k.row_ptr[r]; e < k.row_ptr[r + 1]; ++e)
                    if (rows[order[e]] != r) return 5;
        }
        if (trial == 0) {   // an out-of-range row is reported, not written through
            rows[nnz / 2] = l;
            SketchCsr k;
            k.col_own.assign((size_t)nnz, -1); k.val_own.assign((size_t)nnz, -1.0);
            k.col = k.col_own.data(); k.val = k.val_own.data();
            if (sketch_csr_fill(l, s_eff, ncols, rows.data(), vals.data(), k) != ITCPD_ERR_ARG) return 6;
        }
    }
    // a large case on the default thread count
    unsetenv("ITCPD_SKETCH_THREADS");
    {
        const int l = 5000, s_eff = 4;
        const int64_t ncols = 400000, nnz = ncols * s_eff;
        std::vector<int> rows((size_t)nnz);
        std::vector<double> vals((size_t)nnz);
        for (int64_t q = 0; q < nnz; ++q) { rows[q] = (int)(rnd() %% l); vals[q] = (double)q; }
        SketchCsr k;
        k.col_own.resize((size_t)nnz); k.val_own.resize((size_t)nnz);
        k.col = k.col_own.data(); k.val = k.val_own.data();
        if (sketch_csr_fill(l, s_eff, ncols, rows.data(), vals.data(), k) != ITCPD_OK) return 7;
        for (int r = 0; r < l; ++r)
            for (int64_t e = k.row_ptr[r]; e < k.row_ptr[r + 1]; ++e) {
                const int64_t q = (int64_t)k.val[e];
                if (rows[q] != r || k.col[e] != q / s_eff) return 8;
                if (e > k.row_ptr[r] && k.val[e - 1] >= k.val[e]) return 9;   // increasing non-zero order inside the row
            }
    }
    printf("SKETCH_CSR_OK\n");
    return 0;
}
"""


def test_threaded_counting_sort_is_the_serial_stable_sort(tmp_path):
    src = open(os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "api.cu")).read()
    a = src.index("struct SketchCsr {")
    b = src.index("static int sketch_csr_upload(")
    text = src[a:b]
    assert "static int sketch_csr_fill(" in text and "cuda" not in text.lower().replace("no cuda calls", "")
    cpp = tmp_path / "csr.cpp"
    cpp.write_text(MAIN % text)
    exe = str(tmp_path / "csr")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", str(cpp), "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SKETCH_CSR_OK" in out.stdout, (out.returncode, out.stdout, out.stderr)
