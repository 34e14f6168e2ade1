#!/bin/bash
# Round-2: staged pageable upload (ONE GPU):  gpurun --timeout 900 -- 'bash tools/r2_upload_check.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_edge.py tests/test_gpu_dense.py -m gpu -q -x -k "pageable or als_from_host or golden or smoke" > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log; tail -4 gpurun_out/r2j_tests.log
python - <<'PY'
import time, numpy as np, itcpd
from bench import init_factors
dims, R = (1024, 1024, 1024), 64
T = np.empty(dims, order="F"); T.reshape(-1, order="F")[:] = 0.5
f = init_factors(dims, R)
for staged in (1, 0, 1):
    with itcpd.Engine(0) as e:
        e.set_option("staged_upload", staged)
        t0 = time.perf_counter(); e.set_tensor(T); dt = time.perf_counter() - t0
        t1 = time.perf_counter(); e.set_tensor(T); dt2 = time.perf_counter() - t1
        print(f"staged_upload={staged}: set_tensor(8.59 GB pageable) first {dt:.3f} s ({8.59 / dt:.1f} GB/s), again {dt2:.3f} s ({8.59 / dt2:.1f} GB/s)", flush=True)
    with itcpd.Engine(0) as e:
        e.set_option("staged_upload", staged)
        t0 = time.perf_counter(); e.als_from_host(T, f, 20, dims=dims); dt = time.perf_counter() - t0
        print(f"staged_upload={staged}: als_from_host fresh handle, 20 sweeps: {dt:.3f} s = {20 / dt:.1f} sweeps/s", flush=True)
PY
