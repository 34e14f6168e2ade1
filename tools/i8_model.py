"""Bound model of one contraction pass (dimension-tree GEMM) on one B200: the FP64 DMMA kernel (measured) against the two
INT8 variants of csrc/gemm_i8.cu (not yet run on hardware).  No measurement here: the numbers are the ceilings round 2's
first hardware run is compared with.  Peaks: HBM copy bandwidth from MEASURED_PEAKS.json (fallback 6443 GB/s, the round-1
measurement), FP64 DMMA 37.1 TFLOP/s (profiles/r1_first_light.json), INT8 dense 4.5 POP/s nominal, 128 B/clk/SM shared memory.

  python tools/i8_model.py > profiles/r1_i8_model.txt
"""
import json
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    HBM = 6443.2e9
DMMA, I8, SMS, CLK = 37.1e12, 4.5e15, 148, 1.965e9
SMEM = 128 * CLK * SMS          # B/s, all SMs
BM, BN, BK, NDIG, NACC = 128, 64, 32, 6, 7


def one_pass(rows, kext, R):
    P = rows * kext
    rt, kt, rb = math.ceil(rows / BM), math.ceil(kext / BK), math.ceil(R / BN)
    ksteps = rt * kt * rb
    nq = [NACC - p for p in range(NDIG)]              # B planes per A plane: the 27 digit pairs with p + q <= 6 (7 digits on the Khatri-Rao side)
    mma_ops = ksteps * 2 * BM * BK * BN * sum(nq)     # every rank block padded to 64 columns
    a_bytes, b_reads = NDIG * BM * BK, sum(nq) * BN * BK
    t = {
        "dmma": 2.0 * R * P / DMMA,
        # all rank blocks ride in one launch, neighbouring CTAs stream the same rows of T for different rank blocks: if the L2
        # catches the second reader a pass costs ONE stream of T (shown); if it does not, multiply by the number of rank blocks
        "hbm8": 8.0 * P / HBM,                        # variant 1 streams the FP64 tensor
        "hbm7": 6.0 * P / HBM,                        # variant 2 streams the 6 digit planes
        "mma": mma_ops / I8,
        # shared-memory traffic per k-step: operand reads of the 10 instructions (A planes once, stacked B planes re-read)
        "smem1": ksteps * (a_bytes + b_reads + BM * BK * 8 * 2 + a_bytes + NACC * BN * BK) / SMEM,   # + TMA FP64 in, converter read, planes written, B in
        "smem2": ksteps * (a_bytes + b_reads + a_bytes + NACC * BN * BK) / SMEM,                       # + bulk copies in
        # converters: ~25 integer/logic ops + 1 FP64 mul + 1 F2I per element on 128 int lanes / SM
        "conv": rb * P * 25.0 / (128 * CLK * SMS),
    }
    t["i8_v1"] = max(t["hbm8"], t["mma"], t["smem1"], t["conv"])
    t["i8_v2"] = max(t["hbm7"], t["mma"], t["smem2"])
    return t


CONFIGS = [
    ("A 200^3 R=50 (pass A rows=I0*I1)", 200 * 200, 200, 50),
    ("B 1024^3 R=64", 1024 * 1024, 1024, 64),
    ("C 256^4 R=32 (2,2)", 256 * 256, 256 * 256, 32),
    ("D 2048^3 R=128", 2048 * 2048, 2048, 128),
    ("B slab /8: pass B rows=I1*I2loc", 1024 * 128, 1024, 64),
    ("B slab /8: pass A rows=I0 (split-K)", 1024, 1024 * 128, 64),
    ("D slab /8: pass B", 2048 * 256, 2048, 128),
    ("D slab /8: pass A (split-K)", 2048, 2048 * 256, 128),
]

if __name__ == "__main__":
    print(f"peaks: HBM {HBM / 1e9:.0f} GB/s, DMMA {DMMA / 1e12:.1f} TFLOP/s, INT8 {I8 / 1e15:.1f} POP/s, smem {SMEM / 1e12:.1f} TB/s; times in ms per pass")
    hdr = f"{'config':40s} {'DMMA':>8s} | {'HBM 8B':>8s} {'HBM 6B':>8s} {'MMA i8':>8s} {'smem v1':>8s} {'smem v2':>8s} {'convert':>8s} | {'i8 v1':>8s} {'i8 v2':>8s} {'v2/DMMA':>8s}"
    print(hdr)
    for name, rows, kext, R in CONFIGS:
        t = one_pass(rows, kext, R)
        ms = {k: v * 1e3 for k, v in t.items()}
        print(f"{name:40s} {ms['dmma']:8.3f} | {ms['hbm8']:8.3f} {ms['hbm7']:8.3f} {ms['mma']:8.3f} {ms['smem1']:8.3f} {ms['smem2']:8.3f} {ms['conv']:8.3f} | "
              f"{ms['i8_v1']:8.3f} {ms['i8_v2']:8.3f} {t['dmma'] / t['i8_v2']:8.2f}x")
    print("\nReading: 'i8 v1' / 'i8 v2' are the largest of their bounds (perfect overlap). Variant 1 is bound by the converters' integer")
    print("work and shared-memory traffic, variant 2 by the HBM stream of the digit planes (R <= 64) or by the INT8 pipe (R = 128: two")
    print("rank blocks share one stream of T through the L2, the MMA work doubles).  Config B sweep estimate with variant 2: 2 passes + 0.55 ms of other kernels.")
    t = one_pass(1024 * 1024, 1024, 64)
    print(f"  DMMA today: {1e3 * (2 * 3.92e-3 + 0.55e-3):.2f} ms measured;  variant 2 bound: {1e3 * (2 * t['i8_v2'] + 0.55e-3):.2f} ms -> {1 / (2 * t['i8_v2'] + 0.55e-3):.0f} sweeps/s")
