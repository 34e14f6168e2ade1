#!/usr/bin/env python
"""bench.py -- CP-ALS sweeps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config B] [--no-extras]

A "step" is one ALS sweep (all N mode updates + the fit scalars) of dense random Float64
1024 x 1024 x 1024, rank 64 (BASELINE.json configs[1], "B").  N > 1 (torchrun, one rank per GPU)
slab-shards the SAME tensor along its last mode (strong scaling).

  value   : sweeps/s with the tensor and factors resident in HBM, K sweeps timed with CUDA events on
            the library's stream, max over ranks.
  e2e     : the same metric through the public C-ABI call itcpd_als_from_host with HOST (pinned)
            buffers: H2D of the tensor + factors, K sweeps, D2H of factors/lambda/fit scalars, all
            inside the timed region (one decomposition call; bytes are amortised over its K sweeps).
            `e2e.pageable` repeats the call from a pageable numpy array on a FRESH handle (what a Julia
            Array gives: device allocation and the driver's staging copies inside the timed region).
  roofline: dominant kernel = partial_gemm_kernel (TMA + FP64 DMMA); achieved = 2*R*P flops per launch
            / mean launch time (CUDA events around every launch of a second pass right after the timed region);
            peak = FP64 DMMA issue-rate probe measured live on this GPU with its own clock samples
            (MEASURED_PEAKS.json carries no FP64 number).
  parity  : max |dfit| of every sweep this run executed (warm-up + timed) against the CPU oracle's trajectory on
            the same synthetic tensor and initial factors (tests/golden/bench_trajectory_B.json, written by
            tests/golden/make_bench_trajectory.py) -- at every N, so the scaling records carry multi-GPU parity.
  cpu_baseline: the oracle (numpy/OpenBLAS restatement of the reference, "port") on the host cores: one
            un-extrapolated full-size sweep when host memory allows, else a last-mode slab sample.
  extra   : compact records of the other BASELINE.json configurations: D (2048^3 rank 128, slab-sharded over the
            same N ranks; the whole 68.7 GB tensor at N = 1), and at N = 1 also C (256^4 rank 32), A (200^3
            rank 50, L2 flushed between sweeps) and E (sampled solvers vs exact ALS on a planted tensor).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "A": {"dims": (200, 200, 200), "rank": 50},
    "B": {"dims": (1024, 1024, 1024), "rank": 64},
    "C": {"dims": (256, 256, 256, 256), "rank": 32},
    "D": {"dims": (2048, 2048, 2048), "rank": 128},
    "S": {"dims": (256, 256, 256), "rank": 32},  # small smoke configuration
    "B8": {"dims": (1024, 1024, 128), "rank": 64},  # one rank's slab of config B at 8 GPUs (per-rank chain without collectives)
    "D8": {"dims": (2048, 2048, 256), "rank": 128},  # one rank's slab of config D at 8 GPUs
}


def workload_name(name, dims, R):
    return f"dense random Float64 {'x'.join(map(str, dims))} rank {R} CP-ALS (config {name})"


def init_factors(dims, R, seed=1):
    """randn(I_n, R) from numpy default_rng(seed), column-normalised as src/cpd.jl:48-60."""
    rng = np.random.default_rng(seed)
    out = []
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((I, R)))
        out.append(np.asfortranarray(X / np.sqrt(np.sum(X * X, axis=0))[None, :]))
    return out


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0, period_ms=100):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s in sm if s >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def host_available_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def fits_of(ref_norm, inner, norm2):
    inner, norm2 = np.asarray(inner, dtype=np.float64), np.asarray(norm2, dtype=np.float64)
    return 1.0 - np.sqrt(np.abs(ref_norm * ref_norm + norm2 - 2 * np.abs(inner))) / ref_norm   # fit_check.jl:30-32


def parity_against_golden(name, fits):
    """max |dfit| of the sweeps this run executed (in order, from the initial factors) against a committed trajectory."""
    path = os.path.join(ROOT, "tests", "golden", f"bench_trajectory_{name}.json")
    if not os.path.exists(path):
        return {"reference": None, "note": f"no committed trajectory for config {name}"}
    g = json.load(open(path))
    n = min(len(g["fit"]), len(fits))
    d = np.abs(np.asarray(fits[:n]) - np.asarray(g["fit"][:n]))
    return {"reference": f"tests/golden/bench_trajectory_{name}.json ({g.get('source', '?')})", "sweeps_compared": int(n),
            "max_abs_dfit": float(d.max()) if n else None, "tolerance": 1e-9, "ok": bool(n > 0 and d.max() <= 1e-9),
            "fit_last_compared": float(fits[n - 1]) if n else None}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(Tslab, factors, frac, free_cols=2):
    """One sample of a reference sweep on the host cores over a last-mode slab holding `frac` of the tensor (frac = 1: the
    whole tensor, nothing extrapolated), all N mode updates.
    KRPNormal shape (tensor.jl:12-20): explicit KRP + one GEMM per mode incl. the permuted copy of the
    unfolding; KRPFreeNormal shape (the reference default, tensor.jl:32-44): `free_cols` rank columns of one
    mode, extrapolated to R columns x N modes."""
    from oracle import cpals

    N, R = Tslab.ndim, factors[0].shape[1]
    grams = [cpals.gram(f) for f in factors]
    t0 = time.perf_counter()
    for n in range(N):
        M = cpals.mttkrp_krp_normal(Tslab, factors, n)
        X = cpals.solve_ls_problem(cpals.compute_krp_gram(grams, n), M)
        cpals.row_norm(X)
    t_normal = (time.perf_counter() - t0) / frac
    t_free = None
    if free_cols:
        t1 = time.perf_counter()
        cpals.mttkrp_krp_free(Tslab, factors, 0, ranks=range(free_cols))
        t_free = (time.perf_counter() - t1) / free_cols * R * N / frac
    return t_normal, t_free


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return int(max([p.get("num_threads", 1) for p in threadpool_info()] + [1]))
    except Exception:
        return int(os.cpu_count() or 1)


def slab_factors(factors, slab):
    return factors[:-1] + [np.asfortranarray(factors[-1][:slab, :])]


def cpu_baseline_entry(t_normal, t_free, dims, slab, nsamples):
    whole = slab == dims[-1]
    what = (f"the whole {'x'.join(map(str, dims))} tensor, all mode updates, {nsamples} full sweep(s) timed, nothing extrapolated" if whole else
            f"last-mode slab {'x'.join(map(str, dims[:-1]))}x{slab} (1/{dims[-1] // slab} of the tensor), all mode updates, {nsamples} sample(s), "
            f"scaled by {dims[-1] // slab}")
    return {
        "value": 1.0 / t_normal, "unit": "sweeps/s", "cores": blas_threads(), "kind": "port",
        "sample": f"restated oracle (numpy/OpenBLAS, not Julia), KRPNormal GEMM shape incl. permuted copies: {what} -> {t_normal:.2f} s/sweep"
                  + (f"; reference DEFAULT KRPFreeNormal per-rank loop (2 rank columns of one mode on a 1/8 slab timed, extrapolated to R columns x N modes): {t_free:.1f} s/sweep" if t_free else ""),
        "extrapolated": not whole, "krp_free_default_value": (1.0 / t_free) if t_free else None, "host_cpu_count": os.cpu_count(),
    }


def sample_slab(dims):
    """slab size along the last mode so that one sample is ~1e9 flops-bytes of CPU work (about 1/8 of config B)"""
    last = dims[-1]
    slab = last
    while slab % 2 == 0 and np.prod(dims[:-1]) * slab > (1 << 27):
        slab //= 2
    return slab


# ------------------------------------------------------------------------------------------------
def run_reference(args, cfg):
    """--impl reference: the reference's CPU path (oracle port; Julia is not installed) on the host cores.
    A step is one FULL-SIZE sweep (nothing extrapolated) when the host has the memory for the tensor and its permuted copy;
    `steps` reports the sweeps actually timed (as many of the K asked for as fit in --ref-budget seconds, at least one)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, R = cfg["dims"], cfg["rank"]
    nbytes = float(np.prod(dims)) * 8
    full = host_available_gb() * 2 ** 30 > 2.6 * nbytes + (8 << 30) and not args.ref_slab
    slab = dims[-1] if full else sample_slab(dims)
    sdims = tuple(dims[:-1]) + (slab,)
    t_begin = time.perf_counter()
    rng = np.random.default_rng(0)
    T = np.empty(sdims, order="F")
    flat = T.reshape(-1, order="F")
    for a in range(0, flat.size, 1 << 26):   # in place, chunked: no second copy of an 8.6 GB tensor
        flat[a:a + (1 << 26)] = rng.standard_normal(min(1 << 26, flat.size - a))
    factors = slab_factors(init_factors(dims, R), slab)
    frac = slab / dims[-1]
    nwarm = min(args.warmup, 1) if full else args.warmup
    for _ in range(nwarm):
        cpu_sample(T, factors, frac, free_cols=0)
    tn = []
    t_start = time.perf_counter()
    for i in range(args.steps):
        a, _ = cpu_sample(T, factors, frac, free_cols=0)
        tn.append(a)
        if time.perf_counter() - t_start + a * frac > args.ref_budget:
            break
    timed_wall = time.perf_counter() - t_start
    s8 = max(1, slab // 8)
    _, b = cpu_sample(T[..., :s8], slab_factors(factors, s8), s8 / dims[-1], free_cols=2)
    t_normal = float(np.mean(tn))
    entry = cpu_baseline_entry(t_normal, b, dims, slab, len(tn))
    val = entry["value"]
    line = {"impl": "reference", "metric": "CP-ALS sweeps/sec", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": len(tn), "warmup": nwarm, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, dims, R), "l2": "inputs >> L2"},
            "config_detail": {"steps_requested": args.steps, "timed_wall_s": timed_wall, "total_wall_s": time.perf_counter() - t_begin,
                              "full_size_sweeps": bool(full), "host_available_gb": host_available_gb()},
            "cpu_baseline": entry, "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Job:
    """rendezvous + sharding context of one bench process"""

    def __init__(self, single=False):
        self.world = 1 if single else int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = 0 if single else int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local))
            self.dist, self.torch = dist, torch
        self.comm_ready = False

    def barrier(self, eng):
        eng.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def setup_shards(self, eng, itcpd, dims, R, seed=0):
        """slab `rank` of the global tensor along the last mode, generated on the device; factor slabs; NCCL + peer exchange"""
        world, rank = self.world, self.rank
        assert dims[-1] % world == 0, "last mode must divide evenly over the ranks"
        slab = dims[-1] // world
        ldims = tuple(dims[:-1]) + (slab,)
        stride_last = int(np.prod(dims[:-1]))
        if world > 1 and self.comm_ready:
            eng.peer_disable()   # the exchange buffer is sized for the previous shape
        eng.generate_tensor(ldims, seed=seed, elem_offset=rank * slab * stride_last)
        factors = init_factors(dims, R, seed=1)
        lf = factors[:-1] + [np.asfortranarray(factors[-1][rank * slab:(rank + 1) * slab, :])]
        eng.set_cpd(lf, np.ones(R))
        if world > 1:
            torch, dist = self.torch, self.dist
            if not self.comm_ready:
                uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(itcpd.Engine.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, 0)
                eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
                self.comm_ready = True
            if os.environ.get("ITCPD_PEER", "1") == "1":
                # fused all-reduce + solve over NVLink peer memory (CUDA IPC handles exchanged through torch.distributed)
                if "ITCPD_PEER_GRAPH" in os.environ:   # default on: NCCL-free sweeps, device-side epochs, replayed from a CUDA graph
                    eng.set_option("peer_graph", int(os.environ["ITCPD_PEER_GRAPH"] != "0"))
                mine = torch.frombuffer(bytearray(eng.peer_export()), dtype=torch.uint8).cuda()
                allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.all_gather(allh, mine)
                eng.peer_import(world, rank, b"".join(h.cpu().numpy().tobytes() for h in allh))
        eng.compute_grams()
        return ldims, factors


def timed_sweeps(job, eng, ldims, K, W, sampler=None):
    """W warm-up sweeps, then EXACTLY K sweeps bracketed by barrier + synchronize, CUDA events on the library's stream (the
    sweep body replays a CUDA graph), max over ranks.  Tensors that fit in the L2 get a flush between sweeps."""
    flush = int(np.prod(ldims)) * 8 < (256 << 20)
    W = max(W, 3)
    if flush:   # same call pattern as the timed region, so that the single-sweep graph exists before it
        wi = wn = np.zeros(0)
        for _ in range(W):
            eng.flush_l2()
            eng.sweep_async(1)
    else:
        wi, wn = eng.sweep(W)
    launches0 = eng.launch_count
    if sampler is not None and job.rank == 0:
        sampler.start()   # before the barrier: spawning nvidia-smi must not delay rank 0 inside the other ranks' timed region
    job.barrier(eng)
    eng.event_record(0)
    if flush:
        for _ in range(K):
            eng.flush_l2()
            eng.sweep_async(1)
    else:
        eng.sweep_async(K)
    eng.event_record(1)
    job.barrier(eng)
    ms = job.max_over_ranks(eng.event_elapsed_ms(0, 1))
    clocks = sampler.stop() if (sampler is not None and job.rank == 0) else None
    inner, norm2, fallbacks = eng.sweep_results(1 if flush else K)
    launches = eng.launch_count - launches0
    return {"ms": ms, "K": K, "W": W, "flush": flush, "clocks": clocks, "launches": int(launches), "fallbacks": int(fallbacks),
            "inner": np.concatenate([wi, inner]), "norm2": np.concatenate([wn, norm2])}


def gemm_launch_time(eng, K):
    """CUDA events cannot be recorded inside a graph, so the dominant kernel's launch time is measured right after the timed
    region, same process, same data: events around EVERY GEMM launch of Kr sweeps (stream-K fix-up included)."""
    Kr = max(2, min(K, 10))
    eng.set_option("time_gemm", 1)
    eng.gemm_timing(True)
    eng.sweep_async(Kr)
    gemm_ms, gemm_n = eng.gemm_timing(True)
    eng.set_option("time_gemm", 0)
    return gemm_ms, gemm_n


def probe_fp64_peak(eng, local):
    """FP64 DMMA issue-rate peak (MEASURED_PEAKS.json has none) with its OWN clock samples: the probe is repeated for ~0.4 s
    under a 20 ms nvidia-smi sampler, best repetition kept."""
    s = ClockSampler(local, period_ms=20)
    s.start()
    time.sleep(0.05)
    best, t0, n = 0.0, time.perf_counter(), 0
    while time.perf_counter() - t0 < 0.4 or n < 3:
        best = max(best, eng.probe_dmma_peak())
        n += 1
    return {"dmma_tflops": best, "repetitions": n, "clocks": s.stop()}


def roofline_entry(name, R, P, world, gemm_ms, gemm_n, ms_per_sweep, peak, i8):
    mp = measured_peaks()
    flops_per_launch = 2.0 * R * P / world
    ach = flops_per_launch / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, tnote = None, None
    try:
        t1 = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(name)
        if t1 is not None:
            traffic = float(t1) / world
            tnote = "dram__bytes_read+write per launch from the N=1 `ncu --set full` capture" + (f", scaled by 1/{world} for this rank's slab" if world > 1 else "")
    except Exception:
        pass
    roof = {"bound": "tensor", "kernel": "partial_gemm_kernel (TMA + FP64 DMMA.8x8x4)", "achieved": ach, "peak": peak["dmma_tflops"],
            "unit": "TFLOP/s", "frac": ach / peak["dmma_tflops"], "traffic": traffic, "traffic_source": tnote,
            "peak_source": "FP64 DMMA issue-rate probe measured live in this run (MEASURED_PEAKS.json has no FP64 figure)", "peak_probe": peak,
            "launch_ms": gemm_ms, "launches_timed": gemm_n,
            "launch_timing": "CUDA events around every GEMM launch of a second pass right after the timed region (the timed region replays a CUDA graph)",
            "algorithmic_flops_per_launch": flops_per_launch, "algorithmic_bytes_per_launch": 8.0 * P / world,
            "hbm_achieved_GBs": 8.0 * P / world / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0, "hbm_peak_GBs": mp.get("hbm_gbs"),
            "sweep_roofline_frac": (2 * flops_per_launch / (peak["dmma_tflops"] * 1e12)) / (ms_per_sweep * 1e-3)}
    if i8 and gemm_ms > 0 and mp.get("hbm_gbs"):
        bytes_per_elem = 6.0 if i8 == "2" else 8.0   # pre-packed digit planes stream 6 B per element
        gbs = bytes_per_elem * P / world / (gemm_ms * 1e-3) / 1e9
        roof.update({"bound": "hbm", "kernel": "partial_gemm_i8(p)_kernel (TMA + tcgen05.mma kind::i8 on 6 / 7 base-256 digits, TMEM accumulators)",
                     "achieved": gbs, "peak": mp["hbm_gbs"], "unit": "GB/s", "frac": gbs / mp["hbm_gbs"], "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                     "fp64_equivalent_tflops": ach, "algorithmic_bytes_per_launch": bytes_per_elem * P / world,
                     "sweep_roofline_frac": (2 * bytes_per_elem * P / world / (mp["hbm_gbs"] * 1e9)) / (ms_per_sweep * 1e-3)})
    return roof


def dense_record(job, eng, itcpd, name, K, W, peak, with_parity=True, i8=None):
    """one compact `extra` record: config `name` slab-sharded over the job's ranks (i8: the opt-in INT8 tensor-core contraction)"""
    dims, R = tuple(CONFIGS[name]["dims"]), CONFIGS[name]["rank"]
    P = float(np.prod(dims))
    t0 = time.perf_counter()
    if i8:
        eng.set_option("gemm_i8", int(i8))
    try:
        ldims, _ = job.setup_shards(eng, itcpd, dims, R)
        ref_norm = eng.tensor_norm()
        t = timed_sweeps(job, eng, ldims, K, W)
        gemm_ms, gemm_n = gemm_launch_time(eng, K)
    finally:
        if i8:
            eng.set_option("gemm_i8", 0)
    if job.rank != 0:
        return None
    ms = t["ms"] / K
    roof = roofline_entry(name, R, P, job.world, gemm_ms, gemm_n, ms, peak, i8)
    fits = fits_of(ref_norm, t["inner"], t["norm2"])
    rec = {"config": name if not i8 else f"{name}+gemm_i8={i8}", "workload": workload_name(name, dims, R), "n_gpus": job.world, "value": 1e3 / ms, "unit": "sweeps/s", "ms_per_step": ms,
           "steps": K, "warmup": t["W"], "l2": "flush between steps" if t["flush"] else "inputs >> L2",
           "sharding": f"slab along last mode, {job.world} rank(s)", "gpu_launches": t["launches"], "qrcp_fallbacks": t["fallbacks"],
           "fit_after_timed_sweeps": float(fits[-1]),
           "roofline": {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "launch_ms", "launches_timed", "hbm_achieved_GBs", "sweep_roofline_frac")}}
    if with_parity and not t["flush"]:
        rec["parity"] = parity_against_golden(name, fits)
    if i8:
        rec["dtype"] = "i8 digits (6 x 7 base-256, 48/56-bit fixed point per row) accumulated in int32, FP64 outside the contraction"
        rec["note"] = ("OPT-IN path (option gemm_i8), not the headline: the contraction runs on tcgen05.mma kind::i8 with TMEM accumulators and is "
                       "HBM-bound instead of FP64-pipe-bound; `parity` is against the same FP64 oracle trajectory as the headline")
    rec["seconds"] = time.perf_counter() - t0
    return rec


def config_a_record(eng, itcpd, peak):
    """config A (200^3 rank 50): L2-flushed timing + the full 100-sweep trajectory against the oracle's (north-star bar 1e-9)"""
    rec = dense_record(Job(single=True), eng, itcpd, "A", 100, 3, peak, with_parity=False)
    dims, R = tuple(CONFIGS["A"]["dims"]), CONFIGS["A"]["rank"]
    eng.set_cpd(init_factors(dims, R, seed=1), np.ones(R))
    eng.compute_grams()
    inner, norm2 = eng.sweep(100)
    rec["parity"] = parity_against_golden("A", fits_of(eng.tensor_norm(), inner, norm2))
    return rec


def config_e_record(eng, itcpd, sweeps=20):
    """config E: randomized CP-ALS (leverage-score sampling, SE-QRCS pivot sampling) on 1024^3 rank 64 vs exact ALS, on a planted
    rank-64 + 10 % noise tensor generated on the device (test/rand_cp_als.jl:28-96 asserts sampled fits within 1e-2..1e-1 of exact ALS)."""
    dims, R = tuple(CONFIGS["B"]["dims"]), CONFIGS["B"]["rank"]
    P = float(np.prod(dims))
    noise = 0.1 * np.sqrt(R) / np.sqrt(P)   # noise norm = 10 % of the signal norm
    t_all = time.perf_counter()
    eng.generate_lowrank_tensor(dims, R, seed=11, noise=noise)
    nT = max(eng.tensor_norm(), 1e-300)
    cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))

    def fit_of(cp):
        eng.set_cpd(cp.factors, cp.lam)
        return 1.0 - eng.residual_norm() / nT

    def timed(fn):
        eng.synchronize()
        t0 = time.perf_counter()
        out = fn()
        eng.synchronize()
        return out, time.perf_counter() - t0

    res = []
    chk = itcpd.FitCheck(0.0, sweeps, nT)
    _, dt = timed(lambda: itcpd.als_optimize(eng, cp0, check=chk))
    exact_fit = float(chk.history[-1])
    res.append({"alg": "exact ALS (dimension-tree DMMA path)", "sweeps": sweeps, "ms_per_sweep": 1e3 * dt / sweeps, "fit": exact_fit})
    hbm = measured_peaks().get("hbm_gbs")
    for ns in (10 * R, 64 * R):
        # untimed warm-up (3 sweeps): sizes the scratch buffers for this sample count and captures the sweep graph
        itcpd.als_optimize(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(3), seed=4)
        # set-up (compute_als: factor upload + leverage scores) and sweeps (optimize) timed separately, like the pivot-projected solver below
        als_s, setup_s = timed(lambda: itcpd.compute_als(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(sweeps), seed=5))
        cp, dt = timed(lambda: itcpd.optimize(cp0, als_s))
        nbytes = 3 * 8.0 * (dims[0] * ns + 2 * ns * R + ns * R)   # per sweep: gathered fibres + sampled KRP rows read/written, 8 B per element
        f = float(fit_of(cp))
        res.append({"alg": f"LevScoreSampled({ns})", "setup_s": setup_s, "sweeps": sweeps, "ms_per_sweep": 1e3 * dt / sweeps, "fit": f, "fit_minus_exact": f - exact_fit,
                    "roofline": {"bound": "hbm", "achieved": nbytes / (dt / sweeps) / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": (nbytes / (dt / sweeps) / 1e9 / hbm) if hbm else None, "algorithmic_bytes_per_sweep": nbytes,
                                 "note": "latency-bound: a few MB of gathers per sweep; 70 launches per sweep replayed from one CUDA graph"}})
    ns, ksk = 64 * R, 2 * R
    als, setup = timed(lambda: itcpd.compute_als(eng, cp0, alg=itcpd.SEQRCSPivProjected(1, ns, (1, 2, 3), (ksk,) * 3), check=itcpd.NoCheck(sweeps), seed=9))
    cp, dt = timed(lambda: itcpd.optimize(cp0, als))
    eng.generate_lowrank_tensor(dims, R, seed=11, noise=noise)   # the setup released the dense tensor (like the reference): same seed again for the exact fit
    f = float(fit_of(cp))
    res.append({"alg": f"SEQRCSPivProjected(1,{ns}, rank_vect={ksk})", "setup_s": setup, "sweeps": sweeps, "ms_per_sweep": 1e3 * dt / sweeps, "fit": f,
                "fit_minus_exact": f - exact_fit, "effective_ranks": [int(x) for x in als.additional_items["effective_ranks"]]})
    return {"config": "E", "workload": "randomized CP-ALS on 1024x1024x1024 rank 64 (planted rank-64 tensor + 10 % noise, generated on the device) vs exact ALS fit",
            "n_gpus": 1, "noise_floor_fit": 1.0 - 0.1 / np.sqrt(1.01), "results": res,
            "reference_tolerance": "test/rand_cp_als.jl:43-96: sampled fit within 1e-2 .. 1e-1 of exact ALS",
            "ok": bool(all(r.get("fit_minus_exact", 0.0) > -0.1 for r in res)), "seconds": time.perf_counter() - t_all}


def run_ours(args, cfg):
    import itcpd

    job = Job()
    world, rank, local = job.world, job.rank, job.local
    dims, R = tuple(cfg["dims"]), cfg["rank"]
    K, W = args.steps, args.warmup
    P = float(np.prod(dims))
    eng = itcpd.Engine(local)
    i8 = os.environ.get("ITCPD_GEMM_I8", "0")
    i8 = i8 if i8 in ("1", "2") else None

    # ---- synthetic input: slab `rank` of the global tensor along the last mode, generated on device ----
    ldims, factors = job.setup_shards(eng, itcpd, dims, R)
    ref_norm = eng.tensor_norm()
    peak = probe_fp64_peak(eng, local) if rank == 0 else None

    # ---- warm-up + timed region ----
    t = timed_sweeps(job, eng, ldims, K, W, sampler=ClockSampler(local))
    ms = t["ms"]
    gemm_ms, gemm_n = gemm_launch_time(eng, K)
    phases = None
    if os.environ.get("ITCPD_BENCH_PHASES", "0") == "1":  # diagnostic: where a sweep's time goes (events after every phase, no graph)
        try:
            eng.set_option("time_phases", 1)
            eng.phase_timing(True)
            eng.sweep_async(5)
            ph = eng.phase_timing(True)
            eng.set_option("time_phases", 0)
            phases = {k: v / 5.0 for k, v in ph.items() if k != "marks"}
        except Exception as ex:
            phases = {"error": repr(ex)}
    value = K / (ms * 1e-3)

    line, fits = None, None
    if rank == 0:
        fits = fits_of(ref_norm, t["inner"], t["norm2"])
        roof = roofline_entry(args.config, R, P, world, gemm_ms, gemm_n, ms / K, peak, i8)
        peer = world > 1 and os.environ.get("ITCPD_PEER", "1") == "1"
        shard = f"slab along last mode, {world} rank(s)"
        if peer:
            shard += ", M_n all-reduce fused into the row solve over NVLink peer memory"
            if os.environ.get("ITCPD_PEER_GRAPH", "1") != "0":
                shard += ", sweeps replayed from a CUDA graph (device-side exchange epochs, no NCCL call inside a sweep)"
        line = {"metric": "CP-ALS sweeps/sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": K, "warmup": t["W"],
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64" if not i8 else "i8 digits (6 x 7 base-256, 48/56-bit fixed point per row) accumulated in int32, FP64 outside the contraction",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config, dims, R), "l2": "flush between steps" if t["flush"] else "inputs >> L2"},
                "config_detail": {"algorithm": "normal-equation ALS, two-pass dimension tree, pivoted-Cholesky solve, FitCheck scalars every sweep",
                                  "sharding": shard, "fit_after_timed_sweeps": float(fits[-1]), "qrcp_fallbacks": t["fallbacks"]},
                "parity": parity_against_golden(args.config, fits) if not t["flush"] else None,
                "clocks": t["clocks"], "gpu_launches": t["launches"], "roofline": roof}
        if phases is not None:
            line["config_detail"]["phase_ms_per_sweep"] = phases

    # ---- the README stopping rule, reported separately and outside the timed region: FitCheck(1e-3, 100, |T|)
    # from the same initial guess; on a pure-noise tensor it stops after a few sweeps (fit_check.jl:43-52) ----
    if world == 1:
        try:
            chk = itcpd.FitCheck(1e-3, 100, ref_norm)
            t0 = time.perf_counter()
            itcpd.als_optimize(eng, itcpd.CPD(factors, np.ones(R)), check=chk)
            eng.synchronize()
            line["config_detail"]["readme_rule"] = {"check": "FitCheck(1e-3, 100, norm(T))", "sweeps_to_stop": int(chk.total_iter),
                                                    "final_fit": float(chk.final_fit), "seconds": time.perf_counter() - t0}
        except Exception as ex:
            line["config_detail"]["readme_rule"] = {"error": repr(ex)}

    # ---- end to end through the C-ABI with host buffers (N = 1 only: one call, host tensor) ----
    if world == 1 and not args.no_e2e:
        try:
            pin = itcpd.PinnedBuffer(ldims)
            eng.get_tensor(out=pin)  # fill the host buffer with the same synthetic tensor
            h2d = (8.0 * P + sum(f.size for f in factors) * 8.0) / K
            d2h = (sum(f.size for f in factors) * 8.0 + R * 8.0 + 16.0 * K) / K
            t0 = time.perf_counter()
            fout, lam, inner2, norm22 = eng.als_from_host(pin, factors, K, dims=ldims)
            dt = time.perf_counter() - t0
            line["e2e"] = {"value": K / dt, "unit": "sweeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "call": "itcpd_als_from_host (upload T + factors, K sweeps, download factors/lambda/fit)", "seconds": dt,
                           "sweeps_per_call": K, "host_buffer": "pinned (itcpd_host_alloc), handle warm"}
            fe = fits_of(ref_norm, inner2, norm22)
            m = min(K, len(fits), len(fe))
            line["e2e"]["max_abs_dfit_vs_resident_run"] = float(np.max(np.abs(fe[:m] - fits[:m])))
            if host_available_gb() * 2 ** 30 > 1.3 * 8.0 * P + (8 << 30):
                # what the Julia binding gives: a pageable Array and a handle created for this decomposition (device allocation,
                # the driver's staging copies, first-launch set-up and graph capture all inside the timed region)
                host = np.empty(ldims, order="F")
                np.copyto(host, pin.array)
                t0 = time.perf_counter()
                with itcpd.Engine(local) as e2:
                    e2.als_from_host(host, factors, K, dims=ldims)
                    dt2 = time.perf_counter() - t0
                line["e2e"]["pageable"] = {"value": K / dt2, "unit": "sweeps/s", "seconds": dt2, "host_buffer": "pageable numpy array, fresh handle"}
                del host
            if not args.no_cpu:
                full = host_available_gb() * 2 ** 30 > 1.6 * 8.0 * P + (8 << 30)
                slab = ldims[-1] if full else sample_slab(ldims)
                Ts = pin.array[..., :slab]  # contiguous in column-major order: no copy
                frac = slab / ldims[-1]
                tn = []
                t_start = time.perf_counter()
                while len(tn) < (1 if full else 3) and time.perf_counter() - t_start < args.cpu_budget:
                    a, _ = cpu_sample(Ts, slab_factors(factors, slab), frac, free_cols=0)
                    tn.append(a)
                s8 = max(1, ldims[-1] // 8)
                _, tf = cpu_sample(pin.array[..., :s8], slab_factors(factors, s8), s8 / ldims[-1], free_cols=2)
                line["cpu_baseline"] = cpu_baseline_entry(float(np.mean(tn)), tf, ldims, slab, len(tn))
            pin.free()
        except Exception as ex:  # keep the device-resident number even if the host leg fails
            line.setdefault("e2e", {"value": None, "unit": "sweeps/s"})["error"] = repr(ex)
    elif world > 1 and not args.no_e2e:
        # ---- N > 1: the same call on every rank with ITS slab in pinned host memory (upload slab + factors, K sharded sweeps,
        # download); host wall clock between two barriers, max over ranks.  The attempt is agreed on collectively first: a rank
        # that cannot pin its slab must not leave the others waiting inside a sweep's exchange. ----
        pin, ok = None, 1.0
        try:
            pin = itcpd.PinnedBuffer(ldims)
            eng.get_tensor(out=pin)
        except Exception as ex:
            ok = 0.0
            print(f"[bench rank {rank}] e2e leg skipped: {ex!r}", file=sys.stderr)
        all_ok = -job.max_over_ranks(-ok) > 0.5
        if all_ok:
            lf = factors[:-1] + [np.asfortranarray(factors[-1][rank * ldims[-1]:(rank + 1) * ldims[-1], :])]
            job.barrier(eng)
            t0 = time.perf_counter()
            fout, lam, inner2, norm22 = eng.als_from_host(pin, lf, K, dims=ldims)
            job.barrier(eng)
            dt = job.max_over_ranks(time.perf_counter() - t0)
            if rank == 0:
                fbytes = sum(f.size for f in lf) * 8.0
                line["e2e"] = {"value": K / dt, "unit": "sweeps/s", "h2d_bytes_per_step": (8.0 * P + world * fbytes) / K,
                               "d2h_bytes_per_step": world * (fbytes + R * 8.0 + 16.0 * K) / K,
                               "call": "itcpd_als_from_host on every rank (upload its slab of T + factors, K sharded sweeps, download factors/lambda/fit)",
                               "seconds": dt, "sweeps_per_call": K, "host_buffer": "pinned (itcpd_host_alloc), one slab per rank, handles warm",
                               "timing": "host wall clock between barriers, max over ranks"}
                fe = fits_of(ref_norm, inner2, norm22)
                m = min(K, len(fits), len(fe))
                line["e2e"]["max_abs_dfit_vs_resident_run"] = float(np.max(np.abs(fe[:m] - fits[:m])))
        elif rank == 0:
            line["e2e"] = {"value": None, "unit": "sweeps/s", "error": "a rank could not pin its slab of the tensor in host memory"}
        if pin is not None:
            pin.free()
    elif rank == 0:
        line["e2e"] = {"value": None, "unit": "sweeps/s", "note": "--no-e2e"}

    # ---- the other BASELINE.json configurations, as compact records (each with its own roofline) ----
    if not args.no_extras and args.config == "B":
        extras = []
        watchdog = None
        if rank == 0:
            # the headline must survive a hang in an extra record (a peer that never answers spins for 60 s before it traps): after
            # --extras-timeout seconds rank 0 prints the line it has and leaves
            def bail():
                line["extra"] = extras + [{"config": "watchdog", "error": f"extra records did not finish within {args.extras_timeout} s"}]
                print(json.dumps(line), flush=True)
                os._exit(0)
            watchdog = threading.Timer(args.extras_timeout, bail)
            watchdog.daemon = True
            watchdog.start()

        def attempt(fn, label):
            try:
                r = fn()
                if r is not None:
                    extras.append(r)
            except Exception as ex:
                if rank == 0:
                    extras.append({"config": label, "error": repr(ex)})

        attempt(lambda: dense_record(job, eng, itcpd, "D", 8 if world == 1 else 20, 3, peak), "D")
        if world == 1:
            attempt(lambda: dense_record(job, eng, itcpd, "C", 10, 3, peak), "C")
            attempt(lambda: config_a_record(eng, itcpd, peak), "A")
            attempt(lambda: config_e_record(eng, itcpd), "E")
            attempt(lambda: dense_record(job, eng, itcpd, "B", 20, 3, peak, i8="2"), "B+gemm_i8=2")
        if rank == 0:
            watchdog.cancel()
            line["extra"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if job.dist is not None:
        job.dist.barrier()
        job.dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra records (configs D, C, A, E)")
    ap.add_argument("--extras-timeout", type=float, default=240.0, help="give up on the extra records after this many seconds and print the headline line")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: stop taking sweeps after this many seconds")
    ap.add_argument("--ref-slab", action="store_true", help="reference arm: time a 1/8 last-mode slab and extrapolate (small hosts)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
