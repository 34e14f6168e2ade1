// Host-side sparse-sign embedding generators with the reference's C ABI
//   void f(int l, int n, int s, double* vals, int* rows, int* colstarts)
// (bound by ccall at src/algebra/SEQRCS.jl:41-60).  They consume libc rand() in exactly the order of
// src/algebra/sparse_sign.c:25-70 and src/algebra/sparsestack.c:24-79, so a Julia caller that swaps
// `libsparse` for this library sees bit-identical (vals, rows, colstarts).  Inherently serial
// (one global rand() stream), which is why this piece stays on the host.
#include <cmath>
#include <cstdlib>
#include "../../include/itcpd_b200.h"

namespace {

inline int rand_word_bits() {
    int b = 0;
    for (unsigned long v = (unsigned long)RAND_MAX + 1UL; v > 1UL; v >>= 1) ++b;
    return b;
}

// uniform integer in [0, n) by rejection on rand() (reference: uniform_int)
inline int draw_below(int n) {
    const unsigned long top = (unsigned long)RAND_MAX - (unsigned long)RAND_MAX % (unsigned long)n;
    unsigned long r;
    do { r = (unsigned long)rand(); } while (r > top);
    return (int)(r % (unsigned long)n);
}

struct SignStream {  // bits of rand() words, least significant first
    unsigned int word = 0;
    inline double next(double mag) { const double v = (word & 1u) ? mag : -mag; word >>= 1; return v; }
};

}  // namespace

extern "C" void itcpd_sparse_sign(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    const int zeta = s > l ? l : s;
    const int w = rand_word_bits();
    const double mag = 1.0 / std::sqrt((double)zeta);
    const long nnz = (long)n * zeta;
    SignStream ss;
    ss.word = (unsigned int)rand();
    // full words: note the strict bound -- when nnz is a multiple of the word size the final word's
    // worth of `vals` is left untouched, exactly like the reference (sparse_sign.c:37,45)
    long done = 0;
    for (; done + w < nnz; done += w) {
        for (int b = 0; b < w; ++b) vals[done + b] = ss.next(mag);
        ss.word = (unsigned int)rand();
    }
    for (long q = (long)w * (nnz / w); q < nnz; ++q) vals[q] = ss.next(mag);
    for (int c = 0; c < n + 1; ++c) colstarts[c] = c * zeta;
    for (int c = 0; c < n; ++c) {
        int *col = rows + (long)c * zeta;
        for (int have = 0; have < zeta;) {
            col[have] = draw_below(l);
            bool fresh = true;
            for (int q = 0; q < have && fresh; ++q) fresh = (col[q] != col[have]);
            if (fresh) ++have;
        }
    }
}

extern "C" void itcpd_sparsestack(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    const int zeta = s > l ? l : s;
    const int base = l / zeta, extra = l % zeta;  // the first `extra` row blocks have base+1 rows
    for (int c = 0; c < n + 1; ++c) colstarts[c] = c * zeta;
    const double mag = 1.0 / std::sqrt((double)zeta);
    const int w = rand_word_bits();
    SignStream ss;
    int avail = 0;
    long p = 0;
    for (int c = 0; c < n; ++c) {
        for (int j = 0; j < zeta; ++j, ++p) {
            const int len = base + (j < extra ? 1 : 0);
            const int first = j < extra ? j * (base + 1) : extra * (base + 1) + (j - extra) * base;
            rows[p] = first + draw_below(len);
            if (avail == 0) { ss.word = (unsigned int)rand(); avail = w; }
            vals[p] = ss.next(mag);
            --avail;
        }
    }
}
