"""world_size-2 gloo test (CPU) of the slab-sharded ALS math that the N>1 path implements:
each rank holds T[..., slab] and the slab's rows of the last factor; MTTKRPs of the non-sharded modes are
all-reduced, the sharded mode all-reduces column sums-of-squares, its Gram and the fit inner product.
The collectives here are torch.distributed/gloo; on the GPU the same call sites are NCCL (csrc/comm.cu,
csrc/api.cu: mttkrp_device / gram_device / k_colnorm_scale / k_fit_terms).  The sharded trajectory must equal
the single-process oracle."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allreduce(x: np.ndarray) -> np.ndarray:
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy().reshape(x.shape)


def _worker(rank, world, port, dims, R, nsweeps, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cpals

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(1))
    N = len(dims)
    slab = dims[-1] // world
    sl = slice(rank * slab, (rank + 1) * slab)
    Tl = np.asfortranarray(T[..., sl])
    f = [x.copy() for x in cp.factors[:-1]] + [np.asfortranarray(cp.factors[-1][sl])]
    grams = [cpals.gram(x) for x in f]
    grams[-1] = _allreduce(grams[-1])
    ref_norm = float(np.sqrt(_allreduce(np.array([np.sum(Tl * Tl)]))[0]))
    fits = []
    lam = None
    for _ in range(nsweeps):
        M = None
        for n in range(N):
            G = cpals.compute_krp_gram(grams, n)
            M = cpals.mttkrp_krp_normal(Tl, f, n)
            if n != N - 1:
                M = _allreduce(M)                       # partial sums over slabs
            X = cpals.solve_ls_problem(G, M)
            ss = np.sum(X * X, axis=0)
            if n == N - 1:
                ss = _allreduce(ss)                     # column norms of the sharded factor
            lam = np.sqrt(ss)
            f[n] = np.asfortranarray(X / lam[None, :])
            grams[n] = cpals.gram(f[n])
            if n == N - 1:
                grams[n] = _allreduce(grams[n])
        inner = float(_allreduce(np.array([np.sum(M * (f[-1] * lam[None, :]))]))[0])
        norm2 = cpals.norm_factors(grams, lam)
        fits.append(1.0 - np.sqrt(abs(ref_norm ** 2 + norm2 - 2 * abs(inner))) / ref_norm)
    if rank == 0:
        q.put(fits)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_sharded_als_equals_single_process_oracle():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cpals

    dims, R, nsweeps = (10, 9, 8), 4, 15
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, R, nsweeps, q)) for r in range(2)]
    for p in procs:
        p.start()
    fits = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(1))
    chk = cpals.FitCheck(0.0, nsweeps, float(np.linalg.norm(T)))
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
    assert np.max(np.abs(np.array(fits) - np.array(chk.history))) < 1e-9


def _sampled_worker(rank, world, port, dims, R, nsamp, normal, q):
    """The sharded sampled update of csrc/sampled_sharded.cu, restated with numpy + gloo: owned samples only, K'K and T_s K
    all-reduced (normal=True) or the zero-padded K / T_s summed into full copies (normal=False); the sharded factor solves
    its own rows and all-reduces column norms, Gram and the leverage-score Gram; CPDiff scalars all-reduce the cross-Gram."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cpals, sampled

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(4))
    N = len(dims)
    slab = dims[-1] // world
    off = rank * slab
    Tl = np.asfortranarray(T[..., off:off + slab])
    f = [x.copy() for x in cp.factors[:-1]] + [np.asfortranarray(cp.factors[-1][off:off + slab])]
    prev = [x.copy() for x in f]
    prev_lam = np.ones(R)
    out = {}
    piv_rng = np.random.default_rng(5)          # identical pivots on every rank
    lam = None
    for n in range(N):
        others = [m for m in range(N) if m != n]
        piv = np.stack([piv_rng.integers(1, dims[m] + 1, size=nsamp) for m in others], axis=1)   # GLOBAL 1-based coordinates
        if n != N - 1:
            g = piv[:, -1] - 1 - off
            owned = (g >= 0) & (g < slab)
            lp = piv.copy()
            lp[:, -1] = np.where(owned, g + 1, 1)
            K = sampled.pivot_hadamard([f[m] for m in others], lp) * owned[:, None]
            Ts = sampled.fused_flatten_sample(Tl, n, lp) * owned[None, :]
            if normal:
                Gam = _allreduce(K.T @ K)
                M = _allreduce(Ts @ K)
                X = cpals.solve_ls_problem(Gam, np.asfortranarray(M))
            else:
                X = np.asfortranarray(cpals.ldiv_solve(_allreduce(K), np.asfortranarray(_allreduce(Ts).T)).T)
            ss = np.sum(X * X, axis=0)
        else:
            K = sampled.pivot_hadamard([f[m] for m in others], piv)
            Ts = sampled.fused_flatten_sample(Tl, n, piv)       # my rows of the fibres
            if normal:
                X = cpals.solve_ls_problem(K.T @ K, np.asfortranarray(Ts @ K))
            else:
                X = np.asfortranarray(cpals.ldiv_solve(K, np.asfortranarray(Ts.T)).T)
            ss = _allreduce(np.sum(X * X, axis=0))
        lam = np.sqrt(ss)
        f[n] = np.asfortranarray(X / lam[None, :])
    # leverage scores of the sharded factor from the all-reduced Gram, all-gathered in row order
    G = _allreduce(f[-1].T @ f[-1])
    Linv = np.linalg.inv(np.linalg.cholesky(G))
    lev_loc = np.sum((f[-1] @ Linv.T) ** 2, axis=1) / min(dims[-1], R)
    parts = [torch.zeros(slab, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(lev_loc))
    out["lev_last"] = np.concatenate([p.numpy() for p in parts])
    # CPDiff scalars: the sharded factor's cross-Gram is all-reduced
    cross = [prev[m].T @ f[m] for m in range(N)]
    cross[-1] = _allreduce(cross[-1])
    out["inner_prev_curr"] = float(prev_lam @ np.prod(cross, axis=0) @ lam)
    last = [torch.zeros((slab, R), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(last, torch.from_numpy(np.ascontiguousarray(f[-1])))
    out["factors"] = f[:-1] + [np.concatenate([p.numpy() for p in last], axis=0)]
    out["lam"] = lam
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("normal", [True, False])
def test_slab_sharded_sampled_update_equals_single_process_oracle(normal):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cpals, sampled

    dims, R, nsamp = (9, 8, 10), 3, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampled_worker, args=(r, 2, port, dims, R, nsamp, normal, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process oracle with the same pivots
    rng = np.random.default_rng(3)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(4))
    f = [x.copy() for x in cp.factors]
    prev = [x.copy() for x in f]
    piv_rng = np.random.default_rng(5)
    N = len(dims)
    lam = None
    for n in range(N):
        others = [m for m in range(N) if m != n]
        piv = np.stack([piv_rng.integers(1, dims[m] + 1, size=nsamp) for m in others], axis=1)
        K = sampled.pivot_hadamard([f[m] for m in others], piv)
        Ts = sampled.fused_flatten_sample(T, n, piv)
        if normal:
            X = cpals.solve_ls_problem(K.T @ K, np.asfortranarray(Ts @ K))
        else:
            X = np.asfortranarray(cpals.ldiv_solve(K, np.asfortranarray(Ts.T)).T)
        f[n], lam = cpals.row_norm(X)
    for a, b in zip(got["factors"], f):
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-10
    assert np.linalg.norm(got["lam"] - lam) / np.linalg.norm(lam) < 1e-10
    assert np.linalg.norm(got["lev_last"] - sampled.compute_leverage_score_probability(f[-1])) < 1e-10
    want = float(np.ones(R) @ cpals.cp_cp_inner(prev, f) @ lam)
    assert abs(got["inner_prev_curr"] - want) < 1e-10 * max(1.0, abs(want))
