"""pinned host -> device bandwidth against the NUMA node the pinned pages sit on (is the e2e upload rate a placement effect?)"""
import ctypes, glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
def cpus(path):
    out = []
    for part in open(path).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", [(os.path.basename(n), len(cpus(n + "/cpulist"))) for n in nodes], "affinity now:", len(os.sched_getaffinity(0)), "cpus")
pci = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
for d in glob.glob("/sys/bus/pci/devices/*"):
    try:
        if open(d + "/vendor").read().strip() == "0x10de" and open(d + "/class").read().startswith("0x0302"):
            print("gpu", os.path.basename(d), "numa_node", open(d + "/numa_node").read().strip(), "local_cpulist", open(d + "/local_cpulist").read().strip()[:60])
    except OSError:
        pass
full = os.sched_getaffinity(0)
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def bw(label):
    h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    best = 0
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(h, non_blocking=True); torch.cuda.synchronize()
        best = max(best, (1 << 30) / (time.perf_counter() - t0) / 1e9)
    print(f"{label:30s} {best:6.1f} GB/s", flush=True)
    del h
bw("default placement")
for n in nodes:
    c = [x for x in cpus(n + "/cpulist") if x in full]
    if not c:
        print(os.path.basename(n), "not in the affinity mask"); continue
    os.sched_setaffinity(0, c)
    bw("pinned from " + os.path.basename(n))
os.sched_setaffinity(0, full)
