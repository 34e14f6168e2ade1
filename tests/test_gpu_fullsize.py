"""Full-size (BASELINE.json configs) checks through size-independent properties -- the oracle cannot run at
these sizes in seconds, so the GPU path is checked against itself along independent code paths:
  * every dimension-tree split (kind-0 vs kind-1 GEMM for the same mode) gives the same MTTKRP,
  * rows of the full-size MTTKRP equal the oracle evaluated on a downloaded sub-block (exact same data),
  * the fit identity  ||T - That||^2 = ||T||^2 + ||That||^2 - 2<T,That>  ties fit_terms to the residual kernel,
  * linearity of the MTTKRP in a factor."""
import json
import os

import numpy as np
import pytest

from oracle import cpals

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def init_factors(dims, R, seed=1):
    rng = np.random.default_rng(seed)
    out = []
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((I, R)))
        out.append(np.asfortranarray(X / np.sqrt(np.sum(X * X, axis=0))[None, :]))
    return out


@pytest.fixture(scope="module")
def big(engine):
    info = engine.device_info()
    if info["hbm_bytes"] < 60e9:
        pytest.skip("needs a B200-class memory size")
    dims, R = (1024, 1024, 1024), 64
    engine.set_option("mttkrp_alg", 0)
    engine.generate_tensor(dims, seed=0)
    f = init_factors(dims, R)
    engine.set_cpd(f, np.ones(R))
    engine.compute_grams()
    return engine, dims, R, f


def test_config_b_splits_agree(big):
    eng, dims, R, f = big
    res = {}
    try:
        for sa, sb in [(2, 1), (1, 1), (2, 2)]:
            eng.set_option("split_a", sa)
            eng.set_option("split_b", sb)
            res[(sa, sb)] = [eng.mttkrp(n) for n in range(3)]
    finally:
        eng.set_option("split_a", 0)
        eng.set_option("split_b", 0)
    for n in range(3):
        assert relerr(res[(1, 1)][n], res[(2, 1)][n]) < 1e-12
        assert relerr(res[(2, 2)][n], res[(2, 1)][n]) < 1e-12


def test_config_b_rows_match_oracle_on_downloaded_block(big):
    """M_3[k, :] only involves the slab T[:, :, k]; generate that slab separately (same counter-based stream,
    test_device_generator_statistics_and_slabs proves equality) and evaluate the oracle on it."""
    import itcpd

    eng, dims, R, f = big
    M3 = eng.mttkrp(2)
    with itcpd.Engine(0) as small:
        for k in (0, 517, 1023):
            small.generate_tensor((1024, 1024, 1), seed=0, elem_offset=k * 1024 * 1024)
            slab = small.get_tensor()[:, :, 0]
            ref = np.einsum("ij,ir,jr->r", slab, f[0], f[1])
            assert np.linalg.norm(M3[k] - ref) / np.linalg.norm(ref) < 1e-12


def test_config_b_linearity_and_fit_identity(big):
    eng, dims, R, f = big
    M1 = eng.mttkrp(0)
    g = [x.copy() for x in f]
    g[1] = np.asfortranarray(2.5 * f[1])
    eng.set_factor(1, g[1])
    assert relerr(eng.mttkrp(0), 2.5 * M1) < 1e-13
    eng.set_factor(1, f[1])
    # fit identity after two real sweeps
    eng.compute_grams()
    inner, norm2 = eng.sweep(2)
    nT = eng.tensor_norm()
    resid = eng.residual_norm()
    lhs = resid * resid
    rhs = nT * nT + norm2[-1] - 2 * inner[-1]
    assert abs(lhs - rhs) / (nT * nT) < 1e-12
    assert abs(nT - np.sqrt(2.0 ** 30)) / nT < 1e-3  # i.i.d. N(0,1) entries


def test_config_c_order4_splits_agree(engine):
    """256^4, rank 32 (34 GB): order-4 dimension tree, kind-0 and kind-1 contractions over 65536-long indices."""
    info = engine.device_info()
    if info["hbm_bytes"] < 100e9:
        pytest.skip("needs a B200-class memory size")
    dims, R = (256, 256, 256, 256), 32
    engine.generate_tensor(dims, seed=3)
    f = init_factors(dims, R, seed=4)
    engine.set_cpd(f, np.ones(R))
    res = {}
    try:
        for sa, sb in [(2, 2), (3, 1), (1, 1)]:
            engine.set_option("split_a", sa)
            engine.set_option("split_b", sb)
            res[(sa, sb)] = [engine.mttkrp(n) for n in range(4)]
    finally:
        engine.set_option("split_a", 0)
        engine.set_option("split_b", 0)
    for n in range(4):
        assert relerr(res[(3, 1)][n], res[(2, 2)][n]) < 1e-12
        assert relerr(res[(1, 1)][n], res[(2, 2)][n]) < 1e-12
    engine.compute_grams()
    inner, norm2 = engine.sweep(2)
    nT = engine.tensor_norm()
    resid = engine.residual_norm()
    assert abs(resid * resid - (nT * nT + norm2[-1] - 2 * inner[-1])) / (nT * nT) < 1e-12
    engine.generate_tensor((8, 8, 8), seed=0)  # release the big buffers' contents for the following tests


def _trajectory_vs_golden(engine, name, nsweeps):
    """ALS on the device-generated tensor (seed 0) from bench.py's initial factors against the ORACLE's trajectory on the same
    tensor regenerated on the CPU (tests/golden/make_bench_trajectory.py; oracle/synth.py restates the generator, checked in
    test_gpu_dense.py::test_generator_matches_cpu_restatement).  North-star bar: every sweep's fit within 1e-9."""
    g = json.load(open(os.path.join(GOLDEN, f"bench_trajectory_{name}.json")))
    dims, R = tuple(g["dims"]), g["rank"]
    engine.set_option("mttkrp_alg", 0)
    engine.generate_tensor(dims, seed=0)
    nT = engine.tensor_norm()
    assert abs(nT - g["ref_norm"]) / nT < 1e-13
    engine.set_cpd(init_factors(dims, R), np.ones(R))
    engine.compute_grams()
    inner, norm2 = engine.sweep(nsweeps)
    fits = 1.0 - np.sqrt(np.abs(nT * nT + norm2 - 2 * np.abs(inner))) / nT
    ref = np.array(g["fit"][:nsweeps])
    assert len(ref) == nsweeps
    assert np.max(np.abs(fits - ref)) <= 1e-9, (np.max(np.abs(fits - ref)), int(np.argmax(np.abs(fits - ref))))
    # the two scalars themselves, relative to ||T||^2 (the fit only sees their combination)
    assert np.max(np.abs(inner - np.array(g["inner"][:nsweeps]))) / (nT * nT) < 1e-12
    assert np.max(np.abs(norm2 - np.array(g["norm2"][:nsweeps]))) / (nT * nT) < 1e-12


def test_512_cubed_rank64_trajectory_matches_oracle(engine):
    _trajectory_vs_golden(engine, "B512", 20)


def test_config_b_trajectory_matches_oracle(big):
    """BASELINE.json configs[1] at FULL size: 25 sweeps of 1024^3 rank 64 against the oracle's own sweeps (15 s each on the CPU)."""
    eng, dims, R, f = big
    _trajectory_vs_golden(eng, "B", 25)


def test_config_d_rank128_rows_match_oracle_on_downloaded_blocks(engine):
    """BASELINE.json configs[3] on one GPU at FULL size (2048^3, rank 128: two rank blocks per GEMM, 68.7 GB): rows of all three
    MTTKRPs against numpy on separately generated slabs of the same counter-based stream, both dimension-tree splits, and the
    fit identity after two sweeps."""
    import itcpd

    info = engine.device_info()
    if info["hbm_bytes"] < 150e9:
        pytest.skip("needs the whole 180 GB of a B200")
    dims, R = (2048, 2048, 2048), 128
    I = 2048
    engine.generate_tensor(dims, seed=0)
    f = init_factors(dims, R)
    engine.set_cpd(f, np.ones(R))
    M = [engine.mttkrp(n) for n in range(3)]
    try:
        engine.set_option("split_a", 1)
        engine.set_option("split_b", 1)
        for n in range(3):
            assert relerr(engine.mttkrp(n), M[n]) < 1e-12, n
    finally:
        engine.set_option("split_a", 0)
        engine.set_option("split_b", 0)
    with itcpd.Engine(0) as small:
        for k in (0, 1031, 2047):
            small.generate_tensor((I, I, 1), seed=0, elem_offset=k * I * I)
            slab = small.get_tensor()[:, :, 0]                       # T[:, :, k]
            ref = np.einsum("ij,ir,jr->r", slab, f[0], f[1])
            assert np.linalg.norm(M[2][k] - ref) / np.linalg.norm(ref) < 1e-12, k
        # rows of M_1 and M_2 need T[i, :, :] / T[:, j, :]: assemble them from 2048 single-row reads of a few slabs is too slow;
        # use a 64-slab block instead and compare the block's CONTRIBUTION through linearity in the last factor
        k0, nb = 512, 64
        small.generate_tensor((I, I, nb), seed=0, elem_offset=k0 * I * I)
        blk = small.get_tensor()                                     # T[:, :, k0:k0+nb]
    g = [x.copy() for x in f]
    g[2] = np.zeros_like(f[2])
    g[2][k0:k0 + nb] = f[2][k0:k0 + nb]                              # only the block's rows of the last factor are non-zero
    engine.set_factor(2, np.asfortranarray(g[2]))
    for n in (0, 1):
        ref = cpals.mttkrp_krp_normal(blk, [f[0], f[1], np.asfortranarray(f[2][k0:k0 + nb])], n)
        assert relerr(engine.mttkrp(n), ref) < 1e-12, n
    engine.set_factor(2, f[2])
    engine.compute_grams()
    inner, norm2 = engine.sweep(2)
    nT = engine.tensor_norm()
    resid = engine.residual_norm()
    assert abs(resid * resid - (nT * nT + norm2[-1] - 2 * inner[-1])) / (nT * nT) < 1e-12
    engine.generate_tensor((8, 8, 8), seed=0)  # release the big buffers' contents for the following tests
