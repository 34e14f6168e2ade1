"""Full-size (BASELINE.json configs) checks through size-independent properties -- the oracle cannot run at
these sizes in seconds, so the GPU path is checked against itself along independent code paths:
  * every dimension-tree split (kind-0 vs kind-1 GEMM for the same mode) gives the same MTTKRP,
  * rows of the full-size MTTKRP equal the oracle evaluated on a downloaded sub-block (exact same data),
  * the fit identity  ||T - That||^2 = ||T||^2 + ||That||^2 - 2<T,That>  ties fit_terms to the residual kernel,
  * linearity of the MTTKRP in a factor."""
import numpy as np
import pytest

from oracle import cpals

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
