# Golden vectors from the REAL reference (ITensorCPD.jl itself), for pinning oracle/ and the GPU path (VERDICT r1, item 5a).
#
# The build image has no Julia, so the committed parity evidence is "GPU path == oracle/" with the oracle restating the
# reference.  A maintainer with Julia >= 1.10 closes the loop by running, from a checkout of ITensorCPD.jl:
#
#     julia --project=. /path/to/this/repo/tools/make_julia_golden.jl /path/to/this/repo/tests/golden
#
# which writes tests/golden/julia_dense_als.json, julia_solve.json and julia_sampled.json.  tests/test_julia_golden.py then
# checks the oracle (CPU suite) and the GPU path (-m gpu) against them, with the north-star tolerances (MTTKRP 1e-12 relative
# Frobenius, fit trajectory 1e-9, integer maps / gathers bit-exact); while the files are absent those tests skip and say so.
# Only public entry points and the functions the hot path itself calls are used (file:line in the reference given per block).
using ITensorCPD
using ITensors
using ITensors: NDTensors
using ITensors.NDTensors.Expose: expose
using LinearAlgebra
using Random

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
mkpath(outdir)

# ---- a dependency-free JSON writer (numbers, strings, vectors, dictionaries) ----------------------------------------------
json(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : "null"
json(x::Integer) = string(x)
json(x::Bool) = x ? "true" : "false"
json(x::AbstractString) = "\"" * replace(x, "\\" => "\\\\", "\"" => "\\\"") * "\""
json(x::Nothing) = "null"
json(x::Union{AbstractArray,Tuple}) = "[" * join((json(v) for v in x), ",") * "]"
json(x::AbstractDict) = "{" * join((json(string(k)) * ":" * json(v) for (k, v) in x), ",") * "}"
save(name, d) = open(io -> write(io, json(d)), joinpath(outdir, name), "w")
flat(A::AbstractArray) = vec(collect(Float64, A))            # column-major, first index fastest (SURVEY 8: the layout of every buffer)

# factor matrix of a CPD as I_n x R whatever index order the ITensor stores
fmat(cp, n) = array(cp.factors[n], inds(cp)[n], ITensorCPD.cp_rank(cp))

# ==========================================================================================================================
# 1. dense normal-equation ALS  (decompose.jl:13-30 -> als_optimizer.jl:15-55 -> optimize.jl:6-35; fit_check.jl:24-66)
# ==========================================================================================================================
let
    rng = MersenneTwister(20261018)
    dims, R, nsweeps = (9, 8, 7), 5, 30
    A = randn(rng, dims...)
    is = Index.(dims)
    T = itensor(A, is)
    r = Index(R, "CP rank")
    cp0 = ITensorCPD.random_CPD(T, r; rng = MersenneTwister(3))       # cpd.jl:63-70 (the default stream, made explicit)
    nT = norm(T)

    # MTTKRP of the initial guess under both reference formulations (algorithms/.../standard/tensor.jl:12-44)
    mttkrp = Dict{String,Any}()
    for (name, alg) in (("KRPFreeNormal", ITensorCPD.KRPFreeNormal()), ("KRPNormal", ITensorCPD.KRPNormal()))
        als = ITensorCPD.compute_als(T, cp0; alg, check = ITensorCPD.NoCheck(1))
        mttkrp[name] = [flat(array(ITensorCPD.matricize_tensor(alg, als, cp0.factors, cp0, r, n), is[n], r)) for n in 1:length(dims)]
    end

    # per-sweep fits: one als_optimize call per sweep, fed back (the sweep's state is exactly (factors, lambda): optimize.jl:10-28)
    fits = Float64[]
    cp = cp0
    for s in 1:nsweeps
        chk = ITensorCPD.FitCheck(0.0, 1, nT)
        cp = ITensorCPD.als_optimize(T, cp; alg = ITensorCPD.KRPFreeNormal(), check = chk)
        push!(fits, ITensorCPD.CPDFit(chk))
    end
    # the same run in one call (must end on the same fit), and the README stopping rule (README.md:96-129)
    chk_all = ITensorCPD.FitCheck(0.0, nsweeps, nT)
    cp_all = ITensorCPD.als_optimize(T, cp0; alg = ITensorCPD.KRPFreeNormal(), check = chk_all)
    chk_readme = ITensorCPD.FitCheck(1e-3, 100, nT)
    ITensorCPD.als_optimize(T, cp0; alg = ITensorCPD.KRPFreeNormal(), check = chk_readme)
    rec = ITensorCPD.reconstruct(cp_all)                                  # reconstruct.jl:2-9

    save("julia_dense_als.json", Dict(
        "source" => "ITensorCPD.jl (real reference) via tools/make_julia_golden.jl", "julia" => string(VERSION),
        "dims" => collect(dims), "rank" => R, "T" => flat(A), "ref_norm" => nT,
        "factors0" => [flat(fmat(cp0, n)) for n in 1:length(dims)], "lambda0" => flat(array(cp0.λ)),
        "mttkrp" => mttkrp, "fits" => fits, "final_fit_one_call" => ITensorCPD.CPDFit(chk_all),
        "factors_final" => [flat(fmat(cp_all, n)) for n in 1:length(dims)], "lambda_final" => flat(array(cp_all.λ)),
        "reconstruct_final" => flat(array(rec, is...)),
        "readme_rule" => Dict("total_iter" => chk_readme.total_iter, "final_fit" => chk_readme.final_fit)))
end

# ==========================================================================================================================
# 2. the R x R solve  (MttkrpAlgorithm.jl:34-41 -> ldiv_solve.jl:4-29: pivoted Cholesky tol = 1e-6, pivoted-QR fallback)
# ==========================================================================================================================
let
    rng = MersenneTwister(7)
    cases = Dict{String,Any}[]
    for (label, R, dup) in (("full rank", 6, false), ("rank deficient (two duplicated columns)", 8, true))
        F = [Matrix(qr(randn(rng, 20, R)).Q)[:, 1:R] .+ 0.3 .* randn(rng, 20, R) for _ in 1:2]
        F = [f ./ sqrt.(sum(abs2, f; dims = 1)) for f in F]
        if dup
            for f in F
                f[:, 5] .= f[:, 2]; f[:, 8] .= f[:, 3]
            end
        end
        Gamma = (F[1]' * F[1]) .* (F[2]' * F[2])                          # compute_krp: Hadamard of the Grams (MttkrpAlgorithm.jl:18-31)
        M = randn(rng, 11, R)                                              # an MTTKRP: rows x R
        X = ITensorCPD.ldiv_solve!!(expose(copy(Gamma)), expose(copy(transpose(M))); factorizeA = true)
        push!(cases, Dict("label" => label, "R" => R, "rows" => 11, "Gamma" => flat(Gamma), "M" => flat(M),
                          "X" => flat(copy(transpose(X)))))               # rows x R, as solve_ls_problem returns it (:40)
    end
    save("julia_solve.json", Dict("source" => "ITensorCPD.ldiv_solve!! (ldiv_solve.jl:4-29)", "cases" => cases))
end

# ==========================================================================================================================
# 3. sampled path: leverage scores, index maps, sampled KRP, sampled unfolding, sparse-sign sketch
#    (probability.jl:3-10; pivot_mapping.jl:17-47, :59-85, :111-140; had_contract.jl:277-295; SEQRCS.jl:29-60)
# ==========================================================================================================================
let
    rng = MersenneTwister(11)
    dims, R, nsamp = (12, 10, 8), 4, 9
    A = randn(rng, dims...)
    is = Index.(dims)
    T = itensor(A, is)
    r = Index(R, "CP rank")
    cp = ITensorCPD.random_CPD(T, r; rng = MersenneTwister(5))
    N = length(dims)
    lev = [ITensorCPD.compute_leverage_score_probabilitiy(cp.factors[n], is[n]) for n in 1:N]
    per_mode = Dict{String,Any}[]
    for k in 1:N
        rdims = Tuple(dims[m] for m in 1:N if m != k)
        cols = collect(1:7:prod(rdims))[1:nsamp]                          # a fixed, spread-out list of unfolding columns
        coords = ITensorCPD.column_to_multi_coords(cols, rdims)           # nsamp x (N-1), 1-based
        back = ITensorCPD.multi_coords_to_column(rdims, coords)
        piv = itensor(NDTensors.tensor(NDTensors.Dense(vec(coords)), (Index(nsamp, "pivot"), Index(N - 1))))
        Ts = ITensorCPD.fused_flatten_sample(T, k, piv)                   # dim(k) x nsamp
        K = ITensorCPD.pivot_hadamard([cp.factors[m] for m in 1:N if m != k], r, coords)   # nsamp x R
        # sparse-sign sketch of the mode-k unfolding, matrix-free variant
        n_cols = prod(rdims); m_k = dims[k]
        l = 3 * m_k; s = 2
        vals = Vector{Float64}(undef, n_cols * s); rows = Vector{Int32}(undef, n_cols * s)
        ITensorCPD.sparse_sign_matrix(l, n_cols, s, rows, vals)           # rows come back 1-based (SEQRCS.jl:37)
        Ask = ITensorCPD.sketched_matricization(T, k, l, rows, vals, s)
        push!(per_mode, Dict("mode" => k, "cols" => cols, "coords" => vec(collect(Int, coords)), "cols_back" => collect(Int, back),
                             "gathered" => flat(array(Ts)), "sampled_krp" => flat(array(K)),
                             "sketch" => Dict("l" => l, "s" => s, "rows" => collect(Int, rows), "vals" => vals, "A_sk" => flat(Ask))))
    end
    save("julia_sampled.json", Dict(
        "source" => "ITensorCPD.jl (real reference) via tools/make_julia_golden.jl",
        "dims" => collect(dims), "rank" => R, "nsamp" => nsamp, "T" => flat(A),
        "factors" => [flat(fmat(cp, n)) for n in 1:N], "leverage" => lev, "per_mode" => per_mode))
end

println("wrote julia_dense_als.json, julia_solve.json, julia_sampled.json to ", outdir)
