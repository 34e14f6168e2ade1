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
