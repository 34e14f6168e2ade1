"""Static check of the Julia side of the boundary (Julia is not installed in the build image, so the extension cannot be run):
every `ccall` in julia/ext/ITCPDB200Ext/ITCPDB200Ext.jl and in INTEGRATION.md must name a symbol the shared library exports,
with the argument count and argument kinds (pointer / 32-bit int / 64-bit int / double) of the C header as bound in _lib.py --
a transposed or missing argument in a ccall is a silent memory error at run time, not a compile error.
Also: the extension and the build script resolve the library through ONE shared definition (they disagreed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = os.path.join(ROOT, "itensorcpd.jl_b200", "julia", "ext", "ITCPDB200Ext", "ITCPDB200Ext.jl")
DEPS = os.path.join(ROOT, "itensorcpd.jl_b200", "julia", "deps")

JL_KIND = {"Ptr{Cvoid}": "ptr", "Ref{Ptr{Cvoid}}": "ptr", "Ptr{Float64}": "ptr", "Ref{Float64}": "ptr", "Ptr{Int64}": "ptr", "Ref{Int64}": "ptr",
           "Ptr{UInt8}": "ptr", "Ptr{Int32}": "ptr", "Ptr{Cint}": "ptr", "Ptr{Ptr{Int64}}": "ptr", "Ref{Cint}": "ptr", "Ptr{Ptr{Float64}}": "ptr", "Cstring": "ptr",
           "Cint": "i32", "Int64": "i64", "UInt64": "i64", "Float64": "f64", "Cvoid": "void"}


def c_kind(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int: "i32", C.c_int64: "i64", C.c_uint64: "i64", C.c_double: "f64"}[t]


def ccalls(text):
    pat = re.compile(r"ccall\(\(:(\w+),\s*\w+\),\s*(\w+),\s*\(([^()]*)\)", re.S)
    for m in pat.finditer(text):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip()]
        yield m.group(1), m.group(2), args


def _sigs():
    import itcpd
    return itcpd.package._lib._SIGS


@pytest.mark.parametrize("path", [EXT, os.path.join(ROOT, "INTEGRATION.md")])
def test_every_ccall_matches_the_c_abi(path):
    sigs = _sigs()
    seen = 0
    for sym, ret, args in ccalls(open(path).read()):
        assert sym in sigs, f"{os.path.basename(path)}: ccall of {sym}, which include/itcpd_b200.h does not declare"
        res, argtypes = sigs[sym]
        assert JL_KIND[ret] == c_kind(res), (sym, ret)
        assert len(args) == len(argtypes), f"{sym}: {len(args)} ccall argument types, the C function takes {len(argtypes)}"
        for k, (a, ct) in enumerate(zip(args, argtypes)):
            assert a in JL_KIND, (sym, a)
            assert JL_KIND[a] == c_kind(ct), f"{sym}: argument {k} is {a} in the ccall, {ct} in the C ABI"
        seen += 1
    assert seen >= 20, seen


def test_extension_covers_the_entry_points_the_verdict_asked_for():
    syms = {s for s, _, _ in ccalls(open(EXT).read())}
    need = {"itcpd_create", "itcpd_destroy", "itcpd_set_tensor", "itcpd_set_rank", "itcpd_set_factor", "itcpd_set_lambda", "itcpd_get_factor",
            "itcpd_get_lambda", "itcpd_compute_grams", "itcpd_sweep", "itcpd_leverage_scores", "itcpd_sample_factor_matrices", "itcpd_sampled_update",
            "itcpd_cpd_snapshot", "itcpd_cpd_diff_terms",
            # pivot-projected solvers (qr_lev_score_sampled.jl), reconstruct, multi-GPU init
            "itcpd_qrcp_unfolding", "itcpd_seqrcs_modes", "itcpd_seqrcs_krp", "itcpd_set_projector", "itcpd_projected_update",
            "itcpd_set_shape", "itcpd_reconstruct", "itcpd_residual_norm",
            "itcpd_comm_unique_id", "itcpd_comm_init", "itcpd_peer_export", "itcpd_peer_import", "itcpd_allgather_factor",
            "itcpd_sparse_sign", "itcpd_sparsestack"}
    assert need <= syms, need - syms


def test_library_path_has_one_definition():
    ext, build = open(EXT).read(), open(os.path.join(DEPS, "build_b200.jl")).read()
    paths = open(os.path.join(DEPS, "b200_paths.jl")).read()
    assert 'include(joinpath(@__DIR__, "..", "..", "deps", "b200_paths.jl"))' in ext and "B200Paths.library_path()" in ext
    assert 'include(joinpath(@__DIR__, "b200_paths.jl"))' in build and "B200Paths.build()" in build
    assert "libitcpd_b200.so" not in ext and "libitcpd_b200.so" not in build      # only b200_paths.jl names the file
    assert paths.count('"libitcpd_b200.so"') == 1
    # the development-repository fallback of b200_paths.jl must land on this repository's csrc/ and lib/
    root = os.path.normpath(os.path.join(DEPS, "..", "..", ".."))
    assert os.path.isdir(os.path.join(root, "itensorcpd.jl_b200", "csrc")) and os.path.samefile(root, ROOT)


def test_one_handle_per_device_and_no_per_call_upload():
    """VERDICT f1 / ADVICE: compute_als must not create a handle or upload T per call (rank-adaptive decompose calls it per rank step)"""
    ext = open(EXT).read()
    assert ext.count("Handle(alg.device)") == 1 and "function ITensorCPD.reconstruct" in ext   # the only per-call handle: reconstruct's shape-only scratch handle
    assert ext.count("itcpd_set_tensor") == 1 and "handle_for(target, alg.device)" in ext
    for check in ("CPDiffCheck", "CPAngleCheck", "FitCheck", "NoCheck"):
        assert f"device_check!(check::{check}" in ext, check
