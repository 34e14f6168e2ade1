/* CPU ORACLE (test infrastructure, never linked into the product library).
 *
 * Restatement of the reference's two sparse-sign embedding generators:
 *   sparse_sign()  -> /root/reference/src/algebra/sparse_sign.c:25-70
 *   sparsestack()  -> /root/reference/src/algebra/sparsestack.c:24-79
 * Both draw from libc rand() in a fixed order, so after the same srand() this port must
 * reproduce the reference bit for bit (checked in tests/test_oracle_sparse_sign.py against
 * oracle/_ref/libsparse_sign_ref.so, which oracle/Makefile compiles from the reference
 * sources where they lie).
 *
 * Faithfully kept quirks of the reference:
 *   - sparse_sign writes signs in chunks of 31 (= bits of RAND_MAX+1); when the number of
 *     non-zeros is an exact multiple of 31 the LAST chunk of `vals` is never written
 *     (loop bound `i + 31 < nnz`, sparse_sign.c:37, tail start :45).
 *   - the rejection bound of the uniform draw is `r > limit` with
 *     limit = RAND_MAX - RAND_MAX % n (sparse_sign.c:16-24).
 */
#include <math.h>
#include <stdlib.h>

static int oracle_bits_per_rand(void) {
    long a = (long)RAND_MAX + 1L;
    int b = -1;
    while (a > 0) { a >>= 1; ++b; }
    return b;
}

static int oracle_uniform(int n) {
    const unsigned long lim = (unsigned long)RAND_MAX - ((unsigned long)RAND_MAX % (unsigned long)n);
    int r;
    for (;;) {
        r = rand();
        if ((unsigned long)r <= lim) break;
    }
    return r % n;
}

void oracle_sparse_sign(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    if (s > l) s = l;
    const int bpr = oracle_bits_per_rand();
    const double mag = 1.0 / sqrt((double)s);
    const long nnz = (long)n * s;

    /* signs: one rand() word feeds `bpr` consecutive entries */
    unsigned int word = (unsigned int)rand();
    long base = 0;
    while (base + bpr < nnz) {
        for (long q = base; q < base + bpr; ++q) {
            vals[q] = (word & 1U) ? mag : -mag;
            word >>= 1;
        }
        word = (unsigned int)rand();
        base += bpr;
    }
    for (long q = (long)bpr * (nnz / bpr); q < nnz; ++q) {
        vals[q] = (word & 1U) ? mag : -mag;
        word >>= 1;
    }

    for (int c = 0; c <= n; ++c) colstarts[c] = c * s;

    /* rows: s distinct uniform rows per column, by rejection */
    for (long c0 = 0; c0 < nnz; c0 += s) {
        int have = 0;
        while (have < s) {
            const int cand = oracle_uniform(l);
            rows[c0 + have] = cand;
            int dup = 0;
            for (int q = 0; q < have; ++q)
                if (rows[c0 + q] == cand) { dup = 1; break; }
            if (!dup) ++have;
        }
    }
}

void oracle_sparsestack(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    if (s > l) s = l;
    const int q = l / s, rem = l % s;
    for (int c = 0; c <= n; ++c) colstarts[c] = c * s;
    const double mag = 1.0 / sqrt((double)s);
    const int bpr = oracle_bits_per_rand();
    unsigned int word = 0U;
    int left = 0;
    long p = 0;
    for (int c = 0; c < n; ++c) {
        for (int j = 0; j < s; ++j, ++p) {
            /* rows are split into s near-equal blocks; block j gets one uniform row */
            const int size = (j < rem) ? q + 1 : q;
            const int start = (j < rem) ? j * (q + 1) : rem * (q + 1) + (j - rem) * q;
            rows[p] = start + oracle_uniform(size);
            if (left == 0) { word = (unsigned int)rand(); left = bpr; }
            vals[p] = (word & 1U) ? mag : -mag;
            word >>= 1;
            --left;
        }
    }
}

void oracle_srand(unsigned int seed) { srand(seed); }
