#!/bin/bash
# Round-2 sampled-path check (ONE GPU):  gpurun --timeout 1200 -- 'bash tools/r2_sampled_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampled.py tests/test_julia_golden.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests.log; tail -6 gpurun_out/r2e_tests.log
timeout 300 python tools/r2_sampled_profile.py > gpurun_out/r2e_sampled_timing.jsonl 2> gpurun_out/r2e_err.log; cat gpurun_out/r2e_sampled_timing.jsonl; tail -3 gpurun_out/r2e_err.log
# per-kernel durations + DRAM bytes of the sampled path (ncu: cold-cache, serialised)
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
   -k regex:'gather_fibers|sketch_kernel|qrw_|sampled_mttkrp|pivot_hadamard|cdf_kernel|sample_kernel|quadform|gemm_nn' \
   --log-file gpurun_out/r2e_sampled_ncu.csv python tools/r2_sampled_profile.py > gpurun_out/r2e_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
lines = [l for l in open("gpurun_out/r2e_sampled_ncu.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:48]
    agg[name][r["Metric Name"]] += float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum": cnt[name] += 1
print(f"{'kernel':50s} {'launches':>8s} {'us/launch':>10s} {'MB/launch':>10s} {'GB/s':>8s}")
for name, m in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = cnt[name]; t = m["gpu__time_duration.sum"] / n / 1e3
    b = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n
    print(f"{name:50s} {n:8d} {t:10.1f} {b / 1e6:10.2f} {b / (t * 1e-6) / 1e9 if t > 0 else 0:8.1f}")
PY
