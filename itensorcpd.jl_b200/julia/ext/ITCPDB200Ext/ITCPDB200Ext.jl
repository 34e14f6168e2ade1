# ITCPDB200Ext -- package extension that plugs libitcpd_b200 (hand-written sm_100a CUDA behind a C ABI)
# into ITensorCPD's algorithm-dispatch seam.  It sits next to ext/ITCPDMetalExt (the reference's own
# backend extension, ext/ITCPDMetalExt/ITCPDMetalExt.jl:9-14) and is wired the same way in Project.toml:
#
#   [weakdeps]    Libdl = "8f399da3-3557-5675-b5ff-fb832c97cbdb"
#   [extensions]  ITCPDB200Ext = ["Libdl"]
#
# No CUDA.jl, no ITensors GPU backend: every numeric step is a `ccall` into the shared library, which owns all
# device memory behind an opaque handle.  The public API is untouched:
#
#   using ITensorCPD, Libdl
#   cp = ITensorCPD.decompose(T, 64; alg = ITensorCPD.B200Normal(), check = ITensorCPD.FitCheck(1e-3, 100, norm(T)))
#
# NOTE: Julia is not installed in the build image, so this file could not be executed there; the same control
# flow is exercised through the Python ctypes mirror (itensorcpd.jl_b200/host.py), symbol for symbol.
module ITCPDB200Ext

using ITensorCPD
using ITensorCPD: CPD, ALS, CPDOptimizer, MttkrpAlgorithm, ConvergeAlg, FitCheck, NoCheck, cp_rank, cholesky_epsilon
using ITensors: ITensor, Index, inds, ind, dim, dims, array, itensor, order
using Libdl

# same idiom as the reference's own native helper (src/ITensorCPD.jl:25-36, src/algebra/SEQRCS.jl:41-60): a module global
# holding the library path, assigned in __init__, named directly in the (symbol, library) tuple of every ccall
libitcpd = ""

function __init__()
    lib = joinpath(@__DIR__, "..", "..", "lib", "libitcpd_b200.so")
    isfile(lib) || include(joinpath(@__DIR__, "..", "..", "deps", "build_b200.jl"))
    global libitcpd = lib
end

struct B200Error <: Exception
    code::Cint
    msg::String
end
lasterr() = unsafe_string(ccall((:itcpd_last_error, libitcpd), Cstring, ()))
chk(code) = code == 0 ? nothing : throw(B200Error(code, lasterr()))

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = 0)
        p = Ref{Ptr{Cvoid}}(C_NULL)
        chk(ccall((:itcpd_create, libitcpd), Cint, (Ref{Ptr{Cvoid}}, Cint), p, device))
        h = new(p[])
        finalizer(x -> ccall((:itcpd_destroy, libitcpd), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

## The algorithm object users select: `alg = ITensorCPD.B200Normal()` (normal-equation ALS, like KRPFreeNormal).
## Extensions cannot add names to the parent module, so the 4-line type definition lives in the package itself
## (INTEGRATION.md, patch 1: src/algorithms/als_algorithms/standard/tensor.jl):
##     struct B200Normal <: MttkrpAlgorithm
##         device::Int
##     end
##     B200Normal() = B200Normal(0)
using ITensorCPD: B200Normal

## A CPDOptimizer whose `optimize` drives the library one sweep at a time (als_optimizer.jl:23-24 is duck typed).
struct B200ALS <: CPDOptimizer
    target::ITensor
    mttkrp_alg::B200Normal
    handle::Handle
    check::ConvergeAlg
end

## compute_als hook (optimizers/als_optimizers/standard/tensor.jl:3-14): upload T and the factors, Grams on device.
function ITensorCPD.compute_als(alg::B200Normal, target::ITensor, cp::CPD{<:ITensor};
                                extra_args = Dict(), check = nothing, kwargs...)
    h = Handle(alg.device)
    T = array(target)                       # dense column-major Array{Float64,N}, wrapped without copy
    ds = collect(Int64, size(T))
    chk(ccall((:itcpd_set_tensor, libitcpd), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Float64}), h.ptr, length(ds), ds, T))
    chk(ccall((:itcpd_set_rank, libitcpd), Cint, (Ptr{Cvoid}, Cint), h.ptr, dim(cp_rank(cp))))
    for (n, f) in enumerate(cp.factors)     # ITensor (i_n, r): I_n x R column-major
        chk(ccall((:itcpd_set_factor, libitcpd), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h.ptr, n - 1, array(f)))
    end
    chk(ccall((:itcpd_set_lambda, libitcpd), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, array(cp.λ)))
    chk(ccall((:itcpd_compute_grams, libitcpd), Cint, (Ptr{Cvoid},), h.ptr))
    return B200ALS(target, alg, h, check)
end

## optimize hook (optimizers/als_optimizers/optimize.jl:6-35): the while loop and the convergence state machine stay
## in Julia, each sweep body is one `itcpd_sweep` call returning <T,T̂> and ‖T̂‖² for FitCheck (fit_check.jl:28-29).
function ITensorCPD.optimize(cp::CPD, als::B200ALS; verbose = false)
    h = als.handle.ptr
    rank = cp_rank(cp)
    iter = als.check.iter
    converge = als.check
    inner = Ref{Float64}(0.0); nrm2 = Ref{Float64}(0.0)
    while iter < converge.max_counter
        chk(ccall((:itcpd_sweep, libitcpd), Cint, (Ptr{Cvoid}, Cint, Float64, Ref{Float64}, Ref{Float64}),
                  h, 1, cholesky_epsilon, inner, nrm2))
        b200_check_converge(converge, dim(rank), inner[], nrm2[], verbose) && break
        iter += 1
    end
    factors = Vector{ITensor}()
    for (n, i) in enumerate(inds(cp))
        A = Matrix{Float64}(undef, dim(i), dim(rank))
        chk(ccall((:itcpd_get_factor, libitcpd), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h, n - 1, A))
        push!(factors, itensor(A, i, rank))
    end
    lam = Vector{Float64}(undef, dim(rank))
    chk(ccall((:itcpd_get_lambda, libitcpd), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, lam))
    return CPD{typeof(als.target)}(factors, itensor(lam, rank))
end

## FitCheck fed with the two device scalars: verbatim fit_check.jl:25-65 minus the tensor algebra of :28-29.
function b200_check_converge(check::FitCheck, R, inner_prod, fact_square, verbose)
    check.iter += 1
    normResidual = sqrt(abs(check.ref_norm * check.ref_norm + fact_square - 2 * abs(inner_prod)))
    curr_fit = one(normResidual) - (normResidual / check.ref_norm)
    Δfit = abs(check.lastfit - curr_fit)
    check.lastfit = curr_fit
    verbose && println("$(R)\t $(check.iter) \t $(curr_fit) \t $(Δfit)")
    isnan(curr_fit) && throw("Error NAN")
    if Δfit < check.tolerance
        check.counter += 1
        if check.counter >= 2
            check.total_iter = check.iter; check.iter = 0; check.counter = 0
            check.final_fit = check.lastfit; check.lastfit = 0
            return true
        end
    else
        check.counter = 0
    end
    if check.iter >= check.max_counter
        check.total_iter = check.iter; check.iter = 0; check.counter = 0
        check.final_fit = check.lastfit; check.lastfit = 0
    end
    return false
end
## NoCheck never looks at the factors: no_check.jl:9-20 restated on the rank dimension alone.
function b200_check_converge(check::NoCheck, R, _, __, verbose)
    check.iter += 1
    verbose && println("$(R)\t $(check.iter)")
    if check.iter == check.max_counter
        check.iter = 0
        return true
    end
    return false
end

## ---------------------------------------------------------------------------------------------------------------
## Sampled path: leverage-score sampled ALS with the whole per-mode update on the device.
## Package-side type (INTEGRATION.md patch 1b, next to LevScoreSampled in
## src/algorithms/als_algorithms/randomized/krp_lev_score_sampled.jl:9-17):
##     struct B200LevScoreSampled <: ProjectionAlgorithm
##         NSamples::Tuple
##         device::Int
##     end
##     B200LevScoreSampled(n::Int) = B200LevScoreSampled((n,), 0)
## ---------------------------------------------------------------------------------------------------------------
using ITensorCPD: B200LevScoreSampled, CPDiffCheck, CPAngleCheck

struct B200SampledALS <: CPDOptimizer
    target::ITensor
    mttkrp_alg::B200LevScoreSampled
    handle::Handle
    check::ConvergeAlg
    normal::Bool
    stop_resample::Int
    seed::UInt64
end

## compute_als(::LevScoreSampled) (optimizers/als_optimizers/randomized/krp_lev_score_sampled.jl:1-40): factor weights on device
function ITensorCPD.compute_als(alg::B200LevScoreSampled, target::ITensor, cp::CPD{<:ITensor};
                                extra_args = Dict(), check = nothing, normal = false, stop_resample = -1, seed = 0, kwargs...)
    dense = ITensorCPD.compute_als(B200Normal(alg.device), target, cp; check)   # uploads T, factors, lambda
    h = dense.handle
    for n in 1:length(cp)
        chk(ccall((:itcpd_leverage_scores, libitcpd), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h.ptr, n - 1, C_NULL))
    end
    return B200SampledALS(target, alg, h, check, normal, stop_resample, UInt64(seed))
end

## optimize (optimize.jl:6-35) with the ProjectionAlgorithm hooks (ProjectionAlgorithm.jl:7-68) collapsed into two calls per mode
function ITensorCPD.optimize(cp::CPD, als::B200SampledALS; verbose = false)
    h = als.handle.ptr
    N = length(cp)
    iter = als.check.iter
    pivs = [Matrix{Int64}(undef, (length(als.mttkrp_alg.NSamples) == 1 ? als.mttkrp_alg.NSamples[1] : als.mttkrp_alg.NSamples[n]), N - 1) for n in 1:N]
    drawn = falses(N)
    seed = als.seed
    while iter < als.check.max_counter
        for fact in 1:N
            resample = als.stop_resample < 0 || als.stop_resample > als.check.iter || !drawn[fact]   # krp_lev...:24-27
            if resample
                seed += 1
                chk(ccall((:itcpd_sample_factor_matrices, libitcpd), Cint, (Ptr{Cvoid}, Cint, Int64, UInt64, Ptr{Int64}),
                          h, fact - 1, size(pivs[fact], 1), seed, pivs[fact]))
                drawn[fact] = true
            end
            chk(ccall((:itcpd_sampled_update, libitcpd), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Float64, Cint),
                      h, fact - 1, size(pivs[fact], 1), pivs[fact], cholesky_epsilon, als.normal ? 1 : 0))
        end
        b200_sampled_converged(als.check, h, dim(cp_rank(cp)), verbose) && break
        iter += 1
    end
    return fetch_cpd(h, cp, als.target)
end

## CPDiffCheck from the two device scalars (cp_diff_check.jl:20-71); FitCheck is not supported for sampled solvers
## (ProjectionAlgorithm.jl:30-51): it only counts sweeps.
function b200_sampled_converged(check::CPDiffCheck, h, R, verbose)
    check.iter += 1
    inner = Ref{Float64}(0.0); nrm2 = Ref{Float64}(0.0)
    if isnothing(check.PrevCP)
        chk(ccall((:itcpd_cpd_snapshot, libitcpd), Cint, (Ptr{Cvoid},), h))
        chk(ccall((:itcpd_cpd_diff_terms, libitcpd), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), h, inner, nrm2))
        check.PrevCP = true
        check.norm_prev_iter = nrm2[]
        return false
    end
    chk(ccall((:itcpd_cpd_diff_terms, libitcpd), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), h, inner, nrm2))
    normResidual = sqrt(abs(check.norm_prev_iter + nrm2[] - 2 * abs(inner[])))
    curr_fit = 1.0 - normResidual / sqrt(abs(check.norm_prev_iter))
    Δfit = abs(check.lastfit - curr_fit)
    check.lastfit = curr_fit
    check.norm_prev_iter = nrm2[]
    chk(ccall((:itcpd_cpd_snapshot, libitcpd), Cint, (Ptr{Cvoid},), h))
    verbose && println("$(check.iter) \t $(curr_fit) \t $(Δfit)")
    done = false
    if Δfit < check.tolerance
        check.counter += 1
        done = check.counter >= 2
    else
        check.counter = 0
    end
    if done || check.iter >= check.max_counter
        check.total_iter = check.iter; check.iter = 0; check.counter = 0
        check.final_fit = check.lastfit; check.lastfit = 0; check.PrevCP = nothing
    end
    return done
end
function b200_sampled_converged(check::FitCheck, h, R, verbose)
    check.iter == 0 && println("Warning: FitCheck is not enabled for B200LevScoreSampled will run $(check.max_counter) iterations.")
    check.iter += 1
    check.iter >= check.max_counter && (check.iter = 0)
    return false
end
b200_sampled_converged(check::NoCheck, h, R, verbose) = b200_check_converge(check, R, 0.0, 0.0, verbose)

function fetch_cpd(h, cp::CPD, target)
    rank = cp_rank(cp)
    factors = Vector{ITensor}()
    for (n, i) in enumerate(inds(cp))
        A = Matrix{Float64}(undef, dim(i), dim(rank))
        chk(ccall((:itcpd_get_factor, libitcpd), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h, n - 1, A))
        push!(factors, itensor(A, i, rank))
    end
    lam = Vector{Float64}(undef, dim(rank))
    chk(ccall((:itcpd_get_lambda, libitcpd), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, lam))
    return CPD{typeof(target)}(factors, itensor(lam, rank))
end

## Seam 3: the sparse-sign generators keep the C ABI of libsparse_sign (SEQRCS.jl:41-60); pointing the module
## global `ITensorCPD.libsparse` at libitcpd_b200 and the symbols at itcpd_sparse_sign / itcpd_sparsestack
## yields bit-identical (vals, rows, colstarts).
b200_sparse_sign_call(::Val{false}, l, n, s, vals, rows, colstarts) =
    ccall((:itcpd_sparse_sign, libitcpd), Cvoid, (Cint, Cint, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), l, n, s, vals, rows, colstarts)
b200_sparse_sign_call(::Val{true}, l, n, s, vals, rows, colstarts) =
    ccall((:itcpd_sparsestack, libitcpd), Cvoid, (Cint, Cint, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), l, n, s, vals, rows, colstarts)

end
