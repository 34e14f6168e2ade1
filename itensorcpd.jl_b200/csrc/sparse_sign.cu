// Host-side sparse-sign embedding generators with the reference's C ABI
//   void f(int l, int n, int s, double* vals, int* rows, int* colstarts)
// (bound by ccall at src/algebra/SEQRCS.jl:41-60).  They consume libc rand() in exactly the order of
// src/algebra/sparse_sign.c:25-70 and src/algebra/sparsestack.c:24-79, so a Julia caller that swaps
// `libsparse` for this library sees bit-identical (vals, rows, colstarts).  Inherently serial
// (one global rand() stream), which is why this piece stays on the host.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "../../include/itcpd_b200.h"

namespace {

// ---- the libc stream, two ways ------------------------------------------------------------------------------------------------
// LibcRand calls rand().  On glibc every call takes the generator's lock (about 20 ns; 7 million draws per embedding of a 1024^3
// unfolding = 0.16 s, the critical path of the SE-QRCS set-up once the device work is overlapped).  GlibcStream produces the SAME
// numbers from the SAME global state without the per-call lock: initstate()/setstate() hand out the address of the state array libc
// is using (its first word records the generator type and the position of the rear pointer), the additive-feedback step of glibc's
// TYPE_3 generator (x[f] += x[r]; output x[f] >> 1; 31 words, taps 3 apart -- glibc stdlib/random_r.c) is run on that array in
// place, and setstate() hands it back with the rear pointer's new position, so that later rand() calls -- ours or anybody's --
// continue the stream exactly where plain rand() calls would have left it.  It is used only if a self-test on PRIVATE state buffers
// (the caller's stream is not advanced by it) shows that this libc behaves that way; otherwise, or with ITCPD_PLAIN_RAND set,
// everything goes through rand().
struct LibcRand {
    bool begin() { return true; }
    inline int next() { return rand(); }
    void end() {}
};

class GlibcStream {
    static constexpr int DEG = 31, SEP = 3, TYPES = 5, TYPE_3 = 3;
    int32_t *hdr_ = nullptr;   // libc's state: hdr_[0] = TYPES * rear + type, hdr_[1..31] = the lagged-Fibonacci words
    int f_ = 0, r_ = 0;
    static char *parking() {   // a state for libc to sit on while we hold the real one
        alignas(8) static char buf[128];
        return buf;
    }
    static inline int step(int32_t *x, int &f, int &r) {
        const uint32_t v = (uint32_t)x[f] + (uint32_t)x[r];
        x[f] = (int32_t)v;
        if (++f == DEG) f = 0;
        if (++r == DEG) r = 0;
        return (int)(v >> 1);
    }
    static bool self_test() {
        if (RAND_MAX != 2147483647) return false;
        alignas(8) static char probe[128], copy[128];
        char *global = initstate(1u, parking(), 128);          // libc -> parking; `global` = the caller's stream, untouched below
        if (!global) return false;
        bool ok = false;
        do {
            if (!initstate(20261018u, probe, 128)) break;      // libc -> probe
            for (int q = 0; q < 7; ++q) (void)rand();          // move the pointers off their initial position
            if (setstate(parking()) != probe) break;           // libc -> parking; probe's header now records type and rear
            std::memcpy(copy, probe, 128);
            int32_t *c = reinterpret_cast<int32_t *>(copy);
            if (c[0] % TYPES != TYPE_3) break;
            int r = c[0] / TYPES, f = (r + SEP) % DEG;
            if (r < 0 || r >= DEG) break;
            if (!setstate(probe)) break;                        // libc -> probe again
            bool same = true;
            for (int q = 0; q < 500 && same; ++q) same = rand() == step(c + 1, f, r);
            if (!same) break;
            // hand a state back with an updated header: libc must continue where the in-place steps stopped
            c[0] = TYPES * r + TYPE_3;
            if (setstate(copy) != probe) break;                 // libc -> copy; probe's header refreshed by libc: rear after 500 draws
            int32_t *pr = reinterpret_cast<int32_t *>(probe);
            if (pr[0] != c[0]) break;
            int r2 = pr[0] / TYPES, f2 = (r2 + SEP) % DEG;
            for (int q = 0; q < 100 && same; ++q) same = rand() == step(pr + 1, f2, r2);
            ok = same;
        } while (false);
        setstate(global);                                       // the caller's stream, exactly as it was
        return ok;
    }

public:
    static bool supported() {
        static std::once_flag once;
        static bool ok = false;
        std::call_once(once, [] { ok = getenv("ITCPD_PLAIN_RAND") == nullptr && self_test(); });
        return ok;
    }
    bool begin() {
        if (!supported()) return false;
        hdr_ = reinterpret_cast<int32_t *>(initstate(1u, parking(), 128));   // libc -> parking; the real stream is ours for now
        if (!hdr_) return false;
        r_ = hdr_[0] / TYPES;
        f_ = (r_ + SEP) % DEG;
        if (hdr_[0] % TYPES != TYPE_3 || r_ < 0 || r_ >= DEG) {              // the caller installed another generator type: leave it alone
            setstate(reinterpret_cast<char *>(hdr_));
            hdr_ = nullptr;
            return false;
        }
        return true;
    }
    inline int next() { return step(hdr_ + 1, f_, r_); }
    void end() {
        hdr_[0] = TYPES * r_ + TYPE_3;
        setstate(reinterpret_cast<char *>(hdr_));
        hdr_ = nullptr;
    }
};

inline int rand_word_bits() {
    int b = 0;
    for (unsigned long v = (unsigned long)RAND_MAX + 1UL; v > 1UL; v >>= 1) ++b;
    return b;
}

// uniform integer in [0, n) by rejection on the stream (reference: uniform_int): r = rand() until r <= top, then r % n.
// The modulus is fixed per call site, so the remainder is taken with two multiplications (Lemire's exact 32-bit fastmod:
// r % n == ((M * r mod 2^64) * n) >> 64 with M = floor((2^64 - 1) / n) + 1) instead of a hardware division per draw.
struct Below {
    uint32_t n;
    uint64_t M;
    unsigned long top;
    explicit Below(int n_) : n((uint32_t)n_), M(UINT64_MAX / (uint32_t)n_ + 1), top((unsigned long)RAND_MAX - (unsigned long)RAND_MAX % (unsigned long)n_) {}
    template <class Rng> inline int draw(Rng &rng) const {
        unsigned long r;
        do { r = (unsigned long)rng.next(); } while (r > top);
        const uint64_t low = M * (uint64_t)(uint32_t)r;
        return (int)(((__uint128_t)low * n) >> 64);
    }
};

struct SignStream {  // bits of stream words, least significant first: bit set -> +mag, clear -> -mag
    unsigned int word = 0;
    // the bits are random, so a branch here mispredicts every other value: set the IEEE sign bit directly instead (mag > 0)
    inline double next(double mag) {
        uint64_t b;
        std::memcpy(&b, &mag, 8);
        b |= (uint64_t)(~word & 1u) << 63;
        word >>= 1;
        double v;
        std::memcpy(&v, &b, 8);
        return v;
    }
};

template <class Rng> void sparse_sign_impl(Rng &rng, int l, int n, int s, double *vals, int *rows, int *colstarts) {
    const int zeta = s > l ? l : s;
    const int w = rand_word_bits();
    const double mag = 1.0 / std::sqrt((double)zeta);
    const long nnz = (long)n * zeta;
    SignStream ss;
    ss.word = (unsigned int)rng.next();
    // full words: note the strict bound -- when nnz is a multiple of the word size the final word's
    // worth of `vals` is left untouched, exactly like the reference (sparse_sign.c:37,45)
    long done = 0;
    for (; done + w < nnz; done += w) {
        for (int b = 0; b < w; ++b) vals[done + b] = ss.next(mag);
        ss.word = (unsigned int)rng.next();
    }
    for (long q = (long)w * (nnz / w); q < nnz; ++q) vals[q] = ss.next(mag);
    for (int c = 0; c < n + 1; ++c) colstarts[c] = c * zeta;
    const Below below_l(l);
    for (int c = 0; c < n; ++c) {
        int *col = rows + (long)c * zeta;
        for (int have = 0; have < zeta;) {
            col[have] = below_l.draw(rng);
            bool fresh = true;
            for (int q = 0; q < have && fresh; ++q) fresh = (col[q] != col[have]);
            if (fresh) ++have;
        }
    }
}

template <class Rng> void sparsestack_impl(Rng &rng, int l, int n, int s, double *vals, int *rows, int *colstarts) {
    const int zeta = s > l ? l : s;
    const int base = l / zeta, extra = l % zeta;  // the first `extra` row blocks have base+1 rows
    for (int c = 0; c < n + 1; ++c) colstarts[c] = c * zeta;
    const double mag = 1.0 / std::sqrt((double)zeta);
    const int w = rand_word_bits();
    SignStream ss;
    int avail = 0;
    long p = 0;
    const Below below_long(base + 1), below_short(base > 0 ? base : 1);
    for (int c = 0; c < n; ++c) {
        for (int j = 0; j < zeta; ++j, ++p) {
            const int first = j < extra ? j * (base + 1) : extra * (base + 1) + (j - extra) * base;
            rows[p] = first + (j < extra ? below_long.draw(rng) : below_short.draw(rng));
            if (avail == 0) { ss.word = (unsigned int)rng.next(); avail = w; }
            vals[p] = ss.next(mag);
            --avail;
        }
    }
}

}  // namespace

extern "C" int itcpd_sparse_sign_fast_stream(void) { return GlibcStream::supported() ? 1 : 0; }

extern "C" void itcpd_sparse_sign(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    GlibcStream fast;
    if (fast.begin()) {
        sparse_sign_impl(fast, l, n, s, vals, rows, colstarts);
        fast.end();
    } else {
        LibcRand plain;
        sparse_sign_impl(plain, l, n, s, vals, rows, colstarts);
    }
}

extern "C" void itcpd_sparsestack(int l, int n, int s, double *vals, int *rows, int *colstarts) {
    GlibcStream fast;
    if (fast.begin()) {
        sparsestack_impl(fast, l, n, s, vals, rows, colstarts);
        fast.end();
    } else {
        LibcRand plain;
        sparsestack_impl(plain, l, n, s, vals, rows, colstarts);
    }
}
