"""One line per bench JSON file: sweeps/s, ms per sweep, per-launch GEMM time and its roofline fraction."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f"{path:55s} {d['value']:9.2f} sweeps/s {d['ms_per_step']:8.4f} ms   gemm {r.get('launch_ms', float('nan')):7.4f} ms  {r.get('bound', '?'):6s} frac {r.get('frac', float('nan')):.3f}"
              + (f"  phases {d['config']['phase_ms_per_sweep']}" if "phase_ms_per_sweep" in d.get("config", {}) else ""))
    except Exception as e:  # an empty or truncated file is itself a result (the run died)
        print(f"{path:55s} ERR {e!r}")
