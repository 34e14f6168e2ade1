# Parity of the B200 extension against the package's own CPU path, written in the style of the reference's tests
# (test/cp_als.jl, test/rand_cp_als.jl, test/SEQRCS_test.jl).  For a maintainer with Julia >= 1.10, an sm_100 GPU and the
# INTEGRATION.md patch applied (struct B200 in the package, the extension registered in Project.toml):
#
#     julia --project=. -e 'using Libdl; include("ext/../test/b200_parity.jl")'        (path of this file in the checkout)
#
# NOTE: Julia is not installed in the image this repository is built in; this file has not been executed there.  The same
# comparisons run in Python through the same C symbols (tests/test_gpu_dense.py, tests/test_gpu_sampled.py), with oracle/ standing
# in for the package; this file closes the loop with the package itself, tools/make_julia_golden.jl does so for the fixtures.
using Test, ITensorCPD, ITensors, LinearAlgebra, Random, Libdl
using ITensorCPD: als_optimize, random_CPD, reconstruct, decompose, B200, FitCheck, NoCheck, CPDiffCheck, KRPFreeNormal, KRPNormal,
                  LevScoreSampled, SEQRCSPivProjected, QRPivProjected

relerr(a, b) = norm(a - b) / norm(b)

@testset "B200 dense ALS follows the CPU path sweep by sweep" begin
    i, j, k = Index.((20, 30, 40))
    A = random_itensor(Float64, i, j, k)
    cp0 = random_CPD(A, Index(6, "CP_rank"); rng = MersenneTwister(3))
    nA = norm(A)
    # the state of a sweep is (factors, lambda): feed one-sweep calls back into each other on both paths (optimize.jl:10-28)
    cpu, gpu = cp0, cp0
    for sweep in 1:20
        c1 = FitCheck(0.0, 1, nA); c2 = FitCheck(0.0, 1, nA)
        cpu = als_optimize(A, cpu; alg = KRPFreeNormal(), check = c1)
        gpu = als_optimize(A, gpu; alg = B200(KRPFreeNormal()), check = c2)
        @test abs(ITensorCPD.CPDFit(c1) - ITensorCPD.CPDFit(c2)) <= 1e-9            # north-star bar on the fit trajectory
    end
    @test relerr(reconstruct(gpu), reconstruct(cpu)) < 1e-8
    # one call, the README's stopping rule (README.md:96-129): same number of sweeps, same final fit
    c1 = FitCheck(1e-3, 100, nA); c2 = FitCheck(1e-3, 100, nA)
    als_optimize(A, cp0; alg = KRPNormal(), check = c1)
    als_optimize(A, cp0; alg = B200(KRPNormal()), check = c2)
    @test c1.total_iter == c2.total_iter
    @test abs(c1.final_fit - c2.final_fit) <= 1e-9
end

@testset "B200 decompose / reconstruct, as test/cp_als.jl:9-20" begin
    i, j, k = Index.((20, 30, 40))
    A = random_itensor(Float64, i, j, k)
    opt_A = decompose(A, 400; alg = B200(), check = FitCheck(1e-6, 100, norm(A)))
    @test norm(A - reconstruct(opt_A)) / norm(A) < 1e-5
    @test relerr(reconstruct(opt_A, B200()), reconstruct(opt_A)) < 1e-12              # device reconstruct (reconstruct.jl:2-9)
    # rank-adaptive loop (decompose.jl:32-66): the target is uploaded once, every rank step re-uses the device copy
    low = reconstruct(random_CPD(A, Index(5, "CP_rank"); rng = MersenneTwister(1)))
    opt = decompose(low, 1e-3, 8; alg = B200(), start_rank = 2, rank_step = 3)
    @test norm(low - reconstruct(opt)) / norm(low) < 1e-2
end

@testset "B200 sampled solvers, as test/rand_cp_als.jl" begin
    i, j, k = Index.((30, 35, 25))
    exact = reconstruct(random_CPD(random_itensor(Float64, i, j, k), Index(4, "CP_rank"); rng = MersenneTwister(2)))
    nE = norm(exact)
    cp0 = random_CPD(exact, Index(4, "CP_rank"); rng = MersenneTwister(4))
    ref = als_optimize(exact, cp0; alg = KRPNormal(), check = CPDiffCheck(1e-5, 100))
    e_ref = norm(exact - reconstruct(ref)) / nE
    for alg in (B200(LevScoreSampled(400)),
                B200(QRPivProjected(1, 300)),
                B200(SEQRCSPivProjected(1, 300, (1, 2, 3), (40, 40, 40))))
        ok = false
        for attempt in 1:10                                                         # sampling is random: the reference's tests retry too
            o = als_optimize(exact, cp0; alg, check = CPDiffCheck(1e-5, 100))
            e = norm(exact - reconstruct(o)) / nE
            if e < max(10 * e_ref, 1e-6)
                ok = true
                break
            end
        end
        @test ok
    end
end

@testset "sparse-sign generators: libitcpd_b200 against the package's own C helper, bit for bit" begin
    lib = Base.get_extension(ITensorCPD, :ITCPDB200Ext).libitcpd
    srand(seed) = ccall(:srand, Cvoid, (Cuint,), seed)
    for (sym_ref, sym_b200) in ((:sparse_sign, :itcpd_sparse_sign), (:sparsestack, :itcpd_sparsestack))
        l, n, s = 1200, 10000, 8
        v1 = fill(NaN, n * s); r1 = zeros(Int32, n * s); c1 = zeros(Int32, n + 1)
        v2 = fill(NaN, n * s); r2 = zeros(Int32, n * s); c2 = zeros(Int32, n + 1)
        srand(99)
        ccall(dlsym(dlopen(ITensorCPD.libsparse), sym_ref), Cvoid, (Cint, Cint, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), l, n, s, v1, r1, c1)
        next_ref = ccall(:rand, Cint, ())
        srand(99)
        ccall(dlsym(dlopen(lib), sym_b200), Cvoid, (Cint, Cint, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), l, n, s, v2, r2, c2)
        next_b200 = ccall(:rand, Cint, ())
        @test r1 == r2 && c1 == c2 && isequal(v1, v2)
        @test next_ref == next_b200                                                  # the global rand() stream continues identically
    end
end
