/* The C ABI exercised from C (C99), not through ctypes: includes include/itcpd_b200.h, dlopens libitcpd_b200.so, resolves
 * every entry point it uses by name, and runs a tiny decomposition the way a foreign host (Julia's ccall, cgo, JNI) would:
 * plain pointers and sizes, caller-owned buffers, int status codes, itcpd_last_error for the message.
 *
 *   gcc -std=c99 -Wall -Werror -I include tests/cabi_smoke.c -o cabi_smoke -ldl -lm && ./cabi_smoke itensorcpd.jl_b200/lib/libitcpd_b200.so
 *
 * The tensor is an exact rank-3 tensor, so ALS from a perturbed start must reach fit ~ 1; the MTTKRP returned by the
 * library is also checked against a triple loop.  Test infrastructure (run by tests/test_gpu_cabi_c.py on the GPU box;
 * without a GPU itcpd_create must fail loudly with ITCPD_ERR_NO_DEVICE, which tests/test_cabi_cpu.py checks). */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "itcpd_b200.h"

#define I0 12
#define I1 10
#define I2 8
#define R 3

typedef int (*create_fn)(itcpd_ctx **, int);
typedef int (*destroy_fn)(itcpd_ctx *);
typedef const char *(*err_fn)(void);
typedef int (*set_tensor_fn)(itcpd_ctx *, int, const int64_t *, const double *);
typedef int (*set_rank_fn)(itcpd_ctx *, int);
typedef int (*set_factor_fn)(itcpd_ctx *, int, const double *);
typedef int (*get_factor_fn)(itcpd_ctx *, int, double *);
typedef int (*vec_fn)(itcpd_ctx *, double *);
typedef int (*cvec_fn)(itcpd_ctx *, const double *);
typedef int (*ctx_fn)(itcpd_ctx *);
typedef int (*mttkrp_fn)(itcpd_ctx *, int, double *);
typedef int (*norm_fn)(itcpd_ctx *, double *);
typedef int (*sweep_fn)(itcpd_ctx *, int, double, double *, double *);
typedef int64_t (*count_fn)(itcpd_ctx *);

static void *lib;
static err_fn last_error;

static void *sym(const char *name) {
    void *p = dlsym(lib, name);
    if (!p) { fprintf(stderr, "missing symbol %s\n", name); exit(3); }
    return p;
}

#define CHECK(call)                                                                           \
    do {                                                                                      \
        int st_ = (call);                                                                     \
        if (st_ != ITCPD_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, st_, last_error()); return st_ == ITCPD_ERR_NO_DEVICE ? 77 : 1; } \
    } while (0)

static double frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) / 16777216.0) - 0.5; }

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "itensorcpd.jl_b200/lib/libitcpd_b200.so";
    lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen(%s): %s\n", path, dlerror()); return 2; }
    last_error = (err_fn)sym("itcpd_last_error");
    create_fn create = (create_fn)sym("itcpd_create");
    destroy_fn destroy = (destroy_fn)sym("itcpd_destroy");
    set_tensor_fn set_tensor = (set_tensor_fn)sym("itcpd_set_tensor");
    set_rank_fn set_rank = (set_rank_fn)sym("itcpd_set_rank");
    set_factor_fn set_factor = (set_factor_fn)sym("itcpd_set_factor");
    get_factor_fn get_factor = (get_factor_fn)sym("itcpd_get_factor");
    cvec_fn set_lambda = (cvec_fn)sym("itcpd_set_lambda");
    vec_fn get_lambda = (vec_fn)sym("itcpd_get_lambda");
    ctx_fn compute_grams = (ctx_fn)sym("itcpd_compute_grams");
    mttkrp_fn mttkrp = (mttkrp_fn)sym("itcpd_mttkrp");
    norm_fn tensor_norm = (norm_fn)sym("itcpd_tensor_norm");
    norm_fn residual_norm = (norm_fn)sym("itcpd_residual_norm");
    sweep_fn sweep = (sweep_fn)sym("itcpd_sweep");
    count_fn launch_count = (count_fn)sym("itcpd_launch_count");

    /* exact rank-R tensor, column-major (first index fastest), and a perturbed starting guess with unit columns */
    static double T[I0 * I1 * I2], A0[I0 * R], A1[I1 * R], A2[I2 * R], G0[I0 * R], G1[I1 * R], G2[I2 * R], lam[R], M[I1 * R];
    unsigned seed = 12345u;
    for (int i = 0; i < I0 * R; ++i) A0[i] = frand(&seed);
    for (int i = 0; i < I1 * R; ++i) A1[i] = frand(&seed);
    for (int i = 0; i < I2 * R; ++i) A2[i] = frand(&seed);
    for (int k = 0; k < I2; ++k)
        for (int j = 0; j < I1; ++j)
            for (int i = 0; i < I0; ++i) {
                double v = 0.0;
                for (int r = 0; r < R; ++r) v += A0[i + I0 * r] * A1[j + I1 * r] * A2[k + I2 * r];
                T[i + I0 * (j + I1 * k)] = v;
            }
    double *src[3] = {A0, A1, A2}, *dst[3] = {G0, G1, G2};
    const int ext[3] = {I0, I1, I2};
    for (int n = 0; n < 3; ++n)
        for (int r = 0; r < R; ++r) {
            double s = 0.0;
            for (int i = 0; i < ext[n]; ++i) { dst[n][i + ext[n] * r] = src[n][i + ext[n] * r] + 0.3 * frand(&seed); s += dst[n][i + ext[n] * r] * dst[n][i + ext[n] * r]; }
            for (int i = 0; i < ext[n]; ++i) dst[n][i + ext[n] * r] /= sqrt(s);
        }
    for (int r = 0; r < R; ++r) lam[r] = 1.0;

    itcpd_ctx *ctx = NULL;
    CHECK(create(&ctx, 0));
    const int64_t dims[3] = {I0, I1, I2};
    CHECK(set_tensor(ctx, 3, dims, T));
    CHECK(set_rank(ctx, R));
    for (int n = 0; n < 3; ++n) CHECK(set_factor(ctx, n, dst[n]));
    CHECK(set_lambda(ctx, lam));
    CHECK(compute_grams(ctx));

    /* one hook: the mode-1 MTTKRP against a triple loop (1e-12 relative Frobenius, the north-star bar) */
    CHECK(mttkrp(ctx, 1, M));
    double num = 0.0, den = 0.0;
    for (int r = 0; r < R; ++r)
        for (int j = 0; j < I1; ++j) {
            double v = 0.0;
            for (int k = 0; k < I2; ++k)
                for (int i = 0; i < I0; ++i) v += T[i + I0 * (j + I1 * k)] * G0[i + I0 * r] * G2[k + I2 * r];
            num += (M[j + I1 * r] - v) * (M[j + I1 * r] - v);
            den += v * v;
        }
    if (!(sqrt(num / den) < 1e-12)) { fprintf(stderr, "MTTKRP mismatch: %.3e\n", sqrt(num / den)); return 1; }

    /* the whole loop body 60 times, fit from the two scalars exactly as fit_check.jl:30-32 does */
    double nT = 0.0, inner[60], norm2[60];
    CHECK(tensor_norm(ctx, &nT));
    CHECK(sweep(ctx, 60, 1e-6, inner, norm2));
    const double fit = 1.0 - sqrt(fabs(nT * nT + norm2[59] - 2.0 * fabs(inner[59]))) / nT;
    double resid = 0.0;
    CHECK(residual_norm(ctx, &resid));
    CHECK(get_factor(ctx, 0, G0));
    CHECK(get_lambda(ctx, lam));
    printf("cabi_smoke: MTTKRP rel err %.2e, fit after 60 sweeps %.12f, ||T - That|| / ||T|| %.3e, %lld kernel launches\n", sqrt(num / den), fit,
           resid / nT, (long long)launch_count(ctx));
    if (!(fit > 1.0 - 1e-3) || !(fabs((1.0 - resid / nT) - fit) < 1e-6) || launch_count(ctx) <= 0) { fprintf(stderr, "fit check failed\n"); return 1; }
    CHECK(destroy(ctx));
    dlclose(lib);
    printf("CABI_SMOKE_OK\n");
    return 0;
}
