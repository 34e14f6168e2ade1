"""A CPU stand-in for the C-ABI handle, used ONLY by the `-m "not gpu"` tests to exercise the host-side mirror
(itensorcpd.jl_b200/host.py: optimize loop, convergence state machines, sampled-solver setups, rank-adaptive
decompose) without a GPU.  Every numeric method delegates to the oracle; the product never sees this class."""
import numpy as np

import itcpd
from oracle import cpals, sampled


class FakeEngine(itcpd.Engine):
    def __init__(self, device=0):  # no shared library, no device
        self.device = device
        self.dims = ()
        self.rank = 0
        self.T = None
        self.f, self.lam, self.G, self.M = [], None, [], {}
        self.Gamma = None
        self.X = None
        self.proj = {}
        self.prev = None
        self.options = {}
        self.calls = []

    def close(self):
        pass

    def set_option(self, name, value):
        self.options[name] = value

    # tensor
    def set_tensor(self, T, dims=None):
        self.T = np.asfortranarray(T, dtype=np.float64)
        self.dims = self.T.shape
        self.proj = {}

    def generate_tensor(self, dims, seed=0, elem_offset=0):
        self.T = np.asfortranarray(np.random.default_rng(seed).standard_normal(dims))
        self.dims = tuple(dims)

    def drop_tensor(self):
        self.T = None

    def tensor_norm(self):
        return float(np.linalg.norm(self.T))

    # CPD state
    def set_rank(self, rank):
        self.rank = int(rank)
        self.f = [None] * len(self.dims)
        self.G = [None] * len(self.dims)

    def set_factor(self, mode, A):
        self.f[mode] = np.asfortranarray(A, dtype=np.float64).copy()

    def get_factor(self, mode):
        return self.f[mode].copy(order="F")

    def set_lambda(self, lam):
        self.lam = np.array(lam, dtype=np.float64)

    def get_lambda(self):
        return self.lam.copy()

    def compute_grams(self):
        self.G = [cpals.gram(x) for x in self.f]

    def get_gram(self, mode):
        return self.G[mode].copy()

    # hooks
    def gram_hadamard(self, mode, fetch=True):
        self.Gamma = cpals.compute_krp_gram(self.G, mode)
        return self.Gamma.copy() if fetch else None

    def mttkrp(self, mode, fetch=True):
        assert self.T is not None, "tensor dropped"
        self.M[mode] = cpals.mttkrp_krp_normal(self.T, self.f, mode)
        return self.M[mode].copy() if fetch else None

    def solve(self, mode, chol_tol=1e-6):
        info = {}
        self.X = cpals.solve_ls_problem(self.Gamma, self.M[mode], info)
        return (0 if info["path"] == "cholesky" else 1), info["rank"]

    def normalize(self, mode):
        self.f[mode], self.lam = cpals.row_norm(self.X)

    def post_solve(self, mode):
        self.G[mode] = cpals.gram(self.f[mode])

    def fit_terms(self):
        N = len(self.dims)
        inner = float(np.sum(self.M[N - 1] * (self.f[-1] * self.lam[None, :])))
        return inner, cpals.norm_factors(self.G, self.lam)

    def sweep(self, nsweeps=1, chol_tol=1e-6):
        inner, norm2 = np.empty(nsweeps), np.empty(nsweeps)
        for s in range(nsweeps):
            for n in range(len(self.dims)):
                self.gram_hadamard(n, fetch=False)
                self.mttkrp(n, fetch=False)
                self.solve(n)
                self.normalize(n)
                self.post_solve(n)
            inner[s], norm2[s] = self.fit_terms()
        self.calls.append(("sweep", nsweeps))
        return inner, norm2

    def reconstruct(self):
        return cpals.reconstruct(cpals.CPD(self.f, self.lam))

    # checks
    def cpd_snapshot(self):
        self.prev = ([x.copy() for x in self.f], self.lam.copy())

    def cpd_diff_terms(self):
        pf, pl = self.prev
        return float(pl @ cpals.cp_cp_inner(pf, self.f) @ self.lam), cpals.norm_factors([cpals.gram(x) for x in self.f], self.lam)

    # sampled path
    def leverage_scores(self, mode):
        return sampled.compute_leverage_score_probability(self.f[mode])

    def sample_factor_matrices(self, skip_mode, nsamp, seed):
        probs = [self.leverage_scores(n) for n in range(len(self.dims))]
        return sampled.sample_factor_matrices(nsamp, skip_mode, probs, np.random.default_rng(seed))

    def _ls(self, mode, K, Ts, normal):
        if normal:
            X = cpals.ldiv_solve(K.T @ K, np.asfortranarray((Ts @ K).T))
        else:
            X = cpals.ldiv_solve(K, np.asfortranarray(Ts.T))
        self.f[mode], self.lam = cpals.row_norm(np.asfortranarray(X.T))
        self.G[mode] = cpals.gram(self.f[mode])

    def sampled_update(self, mode, pivots, chol_tol=1e-6, normal=True):
        K = sampled.pivot_hadamard([x for m, x in enumerate(self.f) if m != mode], pivots)
        self._ls(mode, K, sampled.fused_flatten_sample(self.T, mode, pivots), normal)

    def sampled_sweep_async(self, nsweeps, nsamps, draw_counter, chol_tol=1e-6, normal=True):
        """the device-resident sampled sweep: the k-th draw is seeded with draw_counter + k (same seeds as the per-mode calls)"""
        d = int(draw_counter)
        for _ in range(nsweeps):
            for mode in range(len(self.dims)):
                d += 1
                self.sampled_update(mode, self.sample_factor_matrices(mode, nsamps[mode], d), chol_tol, normal)

    def qrcp_unfolding(self, mode):
        _, R, p = sampled.qrcp(cpals.unfold(self.T, mode), want_q=False)
        return p, np.diag(R).copy()

    def qrcp_matrix(self, A, steps=None):
        _, R, p = sampled.qrcp(A, want_q=False)
        return p, np.diag(R).copy()

    def seqrcs(self, mode, l, s, t, injective=False, seed=None):
        info = {}
        _, R, p = sampled.seqrcs_tensor(self.T, mode, l, s, t, injective=injective, seed=seed, info=info)
        return p, np.diag(R).copy(), info["subset"]

    def seqrcs_modes(self, modes, ls, ss, ts, injective=False, seeds=None, use_omega=False):
        return [self.seqrcs(m, l, s, t, injective=injective, seed=None if seeds is None else seeds[i])
                for i, (m, l, s, t) in enumerate(zip(modes, ls, ss, ts))]

    def seqrcs_krp(self, mode, l, s, t, injective=False, seed=None):
        _, R, p = sampled.seqrcs_krp([x for m, x in enumerate(self.f) if m != mode], l, s, t, injective=injective, seed=seed)
        return p, np.diag(R).copy(), 0

    def set_projector(self, mode, pivots):
        piv = np.asfortranarray(pivots, dtype=np.int64)
        self.proj[mode] = (piv, sampled.fused_flatten_sample(self.T, mode, piv))

    def projected_update(self, mode, chol_tol=1e-6, normal=True):
        piv, Ts = self.proj[mode]
        K = sampled.pivot_hadamard([x for m, x in enumerate(self.f) if m != mode], piv)
        self._ls(mode, K, Ts, normal)
