#!/usr/bin/env python
"""bench.py -- CP-ALS sweeps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config B]

A "step" is one ALS sweep (all N mode updates + the fit scalars) of dense random Float64
1024 x 1024 x 1024, rank 64 (BASELINE.json configs[1], "B").  N > 1 (torchrun, one rank per GPU)
slab-shards the SAME tensor along its last mode (strong scaling).

  value   : sweeps/s with the tensor and factors resident in HBM, K sweeps timed with CUDA events on
            the library's stream, max over ranks.
  e2e     : the same metric through the public C-ABI call itcpd_als_from_host with HOST (pinned)
            buffers: H2D of the tensor + factors, K sweeps, D2H of factors/lambda/fit scalars, all
            inside the timed region (one decomposition call; bytes are amortised over its K sweeps).
  roofline: dominant kernel = partial_gemm_kernel (TMA + FP64 DMMA); achieved = 2*R*P flops per launch
            / mean launch time (CUDA events around every launch inside the timed region);
            peak = FP64 DMMA issue-rate probe measured live on this GPU (MEASURED_PEAKS.json carries no
            FP64 number); cuBLAS DGEMM on the same shape is reported beside it.
  cpu_baseline: the oracle (numpy/OpenBLAS restatement of the reference, "port") on the host cores.
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

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_sample(Tslab, factors, frac, free_cols=2):
    """One bounded sample of a reference sweep on the host cores: a last-mode slab holding `frac` of the
    tensor (MTTKRP work is linear in the element count), all N mode updates.
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
    return {
        "value": 1.0 / t_normal, "unit": "sweeps/s", "cores": blas_threads(), "kind": "port",
        "sample": f"restated oracle (numpy/OpenBLAS, not Julia), KRPNormal GEMM shape incl. permuted copies: last-mode slab "
                  f"{'x'.join(map(str, dims[:-1]))}x{slab} (1/{dims[-1] // slab} of the tensor), all mode updates, {nsamples} sample(s), "
                  f"scaled by {dims[-1] // slab} -> {t_normal:.2f} s/sweep"
                  + (f"; reference DEFAULT KRPFreeNormal per-rank loop (2 rank columns sampled, extrapolated): {t_free:.1f} s/sweep" if t_free else ""),
        "krp_free_default_value": (1.0 / t_free) if t_free else None, "host_cpu_count": os.cpu_count(),
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
    Each step is one bounded sample (cpu_sample); W warm-up samples, K timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, R = cfg["dims"], cfg["rank"]
    slab = sample_slab(dims)
    sdims = tuple(dims[:-1]) + (slab,)
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal(sdims))
    factors = slab_factors(init_factors(dims, R), slab)
    frac = slab / dims[-1]
    for _ in range(args.warmup):
        cpu_sample(T, factors, frac, free_cols=0)
    tn, tf = [], []
    t_start = time.perf_counter()
    for i in range(args.steps):
        a, b = cpu_sample(T, factors, frac, free_cols=2 if i == 0 else 0)
        tn.append(a)
        if b:
            tf.append(b)
        if time.perf_counter() - t_start > args.ref_budget:
            break
    t_normal = float(np.mean(tn))
    entry = cpu_baseline_entry(t_normal, tf[0] if tf else None, dims, slab, len(tn))
    val = entry["value"]
    line = {"impl": "reference", "metric": "CP-ALS sweeps/sec", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": len(tn), "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"dense random Float64 {'x'.join(map(str, dims))} rank {R} CP-ALS (config {args.config})"},
            "cpu_baseline": entry, "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, cfg):
    import itcpd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dims, R = tuple(cfg["dims"]), cfg["rank"]
    N = len(dims)
    K, W = args.steps, args.warmup
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    eng = itcpd.Engine(local)
    info = eng.device_info()

    # ---- synthetic input: slab `rank` of the global tensor along the last mode, generated on device ----
    assert dims[-1] % world == 0, "last mode must divide evenly over the ranks"
    slab = dims[-1] // world
    ldims = dims[:-1] + (slab,)
    P = float(np.prod(dims))
    stride_last = int(np.prod(dims[:-1]))
    eng.generate_tensor(ldims, seed=0, elem_offset=rank * slab * stride_last)
    factors = init_factors(dims, R, seed=1)
    lf = factors[:-1] + [np.asfortranarray(factors[-1][rank * slab:(rank + 1) * slab, :])]
    eng.set_cpd(lf, np.ones(R))
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(itcpd.Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
        if os.environ.get("ITCPD_PEER", "1") == "1":
            # fused all-reduce + solve over NVLink peer memory (CUDA IPC handles exchanged through torch.distributed)
            if "ITCPD_PEER_GRAPH" in os.environ:   # default on: NCCL-free sweeps with device-side epochs, replayed from a CUDA graph
                eng.set_option("peer_graph", int(os.environ["ITCPD_PEER_GRAPH"] != "0"))
            mine = torch.frombuffer(bytearray(eng.peer_export()), dtype=torch.uint8).cuda()
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            eng.peer_import(world, rank, b"".join(h.cpu().numpy().tobytes() for h in allh))
    eng.compute_grams()
    ref_norm = eng.tensor_norm()

    def barrier():
        eng.synchronize()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    flush = int(np.prod(ldims)) * 8 < (256 << 20)  # tensors that fit in the 126 MB L2 get a flush between steps

    # ---- warm-up ----
    eng.sweep(max(W, 3))
    peaks = {}
    if rank == 0:
        peaks["dmma_tflops"] = eng.probe_dmma_peak()
    # ---- timed region: exactly K sweeps, CUDA events on the library's stream (the sweep body replays a CUDA graph) ----
    launches0 = eng.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # before the barrier: spawning nvidia-smi must not delay rank 0 inside the other ranks' timed region
    barrier()
    eng.event_record(0)
    if flush:
        for _ in range(K):
            eng.flush_l2()
            eng.sweep_async(1)
    else:
        eng.sweep_async(K)
    eng.event_record(1)
    barrier()
    ms = eng.event_elapsed_ms(0, 1)
    clocks = sampler.stop() if rank == 0 else None
    inner, norm2, fallbacks = eng.sweep_results(1 if flush else K)
    launches = eng.launch_count - launches0
    # ---- roofline pass: CUDA events cannot be recorded inside a graph, so the dominant kernel's launch time is
    # measured right after the timed region, same process, same data: events around EVERY GEMM launch of Kr sweeps ----
    Kr = max(2, min(K, 10))
    eng.set_option("time_gemm", 1)
    eng.gemm_timing(True)
    eng.sweep_async(Kr)
    gemm_ms, gemm_n = eng.gemm_timing(True)
    eng.set_option("time_gemm", 0)
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
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = K / (ms * 1e-3)
    fit_last = 1.0 - np.sqrt(abs(ref_norm ** 2 + norm2[-1] - 2 * abs(inner[-1]))) / ref_norm

    line = None
    if rank == 0:
        mp = measured_peaks()
        flops_per_launch = 2.0 * R * P / world
        ach = flops_per_launch / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(args.config)
        except Exception:
            pass
        i8 = os.environ.get("ITCPD_GEMM_I8", "0") in ("1", "2")   # experimental INT8 tensor-core contraction: the pass is then HBM-bound
        roof = {"bound": "tensor", "kernel": "partial_gemm_kernel (TMA + FP64 DMMA.8x8x4)", "achieved": ach, "peak": peaks["dmma_tflops"],
                "unit": "TFLOP/s", "frac": ach / peaks["dmma_tflops"], "traffic": traffic,
                "peak_source": "FP64 DMMA issue-rate probe measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "launch_ms": gemm_ms, "launches_timed": gemm_n,
                "launch_timing": "CUDA events around every GEMM launch of a second pass right after the timed region (the timed region replays a CUDA graph)",
                "algorithmic_flops_per_launch": flops_per_launch,
                "algorithmic_bytes_per_launch": 8.0 * P / world,
                "hbm_achieved_GBs": 8.0 * P / world / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0, "hbm_peak_GBs": mp.get("hbm_gbs"),
                "sweep_roofline_frac": (2 * flops_per_launch / (peaks["dmma_tflops"] * 1e12)) / (ms * 1e-3 / K)}
        if i8 and gemm_ms > 0 and mp.get("hbm_gbs"):
            bytes_per_elem = 6.0 if os.environ.get("ITCPD_GEMM_I8") == "2" else 8.0   # pre-packed digit planes stream 6 B per element
            gbs = bytes_per_elem * P / world / (gemm_ms * 1e-3) / 1e9
            roof.update({"bound": "hbm", "kernel": "partial_gemm_i8(p)_kernel (TMA + tcgen05.mma kind::i8 on 6 / 7 base-256 digits, TMEM accumulators)",
                         "achieved": gbs, "peak": mp["hbm_gbs"], "unit": "GB/s", "frac": gbs / mp["hbm_gbs"],
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs", "fp64_equivalent_tflops": ach,
                         "algorithmic_bytes_per_launch": bytes_per_elem * P / world,
                         "sweep_roofline_frac": (2 * bytes_per_elem * P / world / (mp["hbm_gbs"] * 1e9)) / (ms * 1e-3 / K)})
        line = {"metric": "CP-ALS sweeps/sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"dense random Float64 {'x'.join(map(str, dims))} rank {R} CP-ALS (config {args.config})",
                           "algorithm": "normal-equation ALS, two-pass dimension tree, pivoted-Cholesky solve, FitCheck scalars every sweep",
                           "sharding": f"slab along last mode, {world} rank(s)" + (", M_n all-reduce fused into the row solve over NVLink peer memory"
                                                                                       if world > 1 and os.environ.get("ITCPD_PEER", "1") == "1" else ""), "l2": "flush between steps" if flush else "inputs >> L2",
                           "fit_after_timed_sweeps": float(fit_last), "qrcp_fallbacks": int(fallbacks)},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roof}
        if phases is not None:
            line["config"]["phase_ms_per_sweep"] = phases

    # ---- the README stopping rule, reported separately and outside the timed region: FitCheck(1e-3, 100, |T|)
    # from the same initial guess; on a pure-noise tensor it stops after a few sweeps (fit_check.jl:43-52) ----
    if world == 1:
        try:
            chk = itcpd.FitCheck(1e-3, 100, ref_norm)
            t0 = time.perf_counter()
            itcpd.als_optimize(eng, itcpd.CPD(factors, np.ones(R)), check=chk)
            eng.synchronize()
            line["config"]["readme_rule"] = {"check": "FitCheck(1e-3, 100, norm(T))", "sweeps_to_stop": int(chk.total_iter),
                                             "final_fit": float(chk.final_fit), "seconds": time.perf_counter() - t0}
        except Exception as ex:
            line["config"]["readme_rule"] = {"error": repr(ex)}

    # ---- end to end through the C-ABI with host buffers (N = 1 only: one call, host tensor) ----
    if world == 1 and not args.no_e2e:
        try:
            pin = itcpd.PinnedBuffer(ldims)
            eng.get_tensor(out=pin)  # fill the host buffer with the same synthetic tensor
            t0 = time.perf_counter()
            fout, lam, inner2, norm22 = eng.als_from_host(pin, factors, K, dims=ldims)
            dt = time.perf_counter() - t0
            h2d = (8.0 * P + sum(f.size for f in factors) * 8.0) / K
            d2h = (sum(f.size for f in factors) * 8.0 + R * 8.0 + 16.0 * K) / K
            line["e2e"] = {"value": K / dt, "unit": "sweeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "call": "itcpd_als_from_host (upload T + factors, K sweeps, download factors/lambda/fit)", "seconds": dt,
                           "sweeps_per_call": K}
            if not args.no_cpu:
                slab = sample_slab(ldims)
                Ts = pin.array[..., :slab]  # contiguous in column-major order: no copy
                frac = slab / ldims[-1]
                tn, tf = [], None
                t_start = time.perf_counter()
                while len(tn) < 3 and time.perf_counter() - t_start < args.cpu_budget:
                    a, b = cpu_sample(Ts, slab_factors(factors, slab), frac, free_cols=2 if not tn else 0)
                    tn.append(a)
                    tf = tf or b
                line["cpu_baseline"] = cpu_baseline_entry(float(np.mean(tn)), tf, ldims, slab, len(tn))
            pin.free()
        except Exception as ex:  # keep the device-resident number even if the host leg fails
            line["e2e"] = {"value": None, "unit": "sweeps/s", "error": repr(ex)}
    elif rank == 0:
        line["e2e"] = {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "multi-rank run: e2e is measured at N=1 only"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra records (configs A, C, D, E)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: stop taking samples after this many seconds")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
