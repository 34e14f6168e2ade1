"""Run under torchrun (one rank per GPU): slab-sharded ALS through the C-ABI + NCCL vs the single-GPU result.
Prints 'MULTI_GPU_OK' on rank 0 when the fit trajectories agree to 1e-10 and the gathered factors to 1e-8."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import itcpd


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims, R, nsweeps = (64, 48, 8 * world), 12, 12
    rng = np.random.default_rng(1)
    factors = []
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((I, R)))
        factors.append(np.asfortranarray(X / np.sqrt(np.sum(X * X, axis=0))[None, :]))
    slab = dims[-1] // world
    eng = itcpd.Engine(local)
    eng.set_option("use_graph", int(os.environ.get("ITCPD_GRAPH", "1")))
    eng.generate_tensor(dims[:-1] + (slab,), seed=7, elem_offset=rank * slab * dims[0] * dims[1])
    eng.set_cpd(factors[:-1] + [np.asfortranarray(factors[-1][rank * slab:(rank + 1) * slab])], np.ones(R))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(itcpd.Engine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    eng.comm_init(world, rank, uid.cpu().numpy().tobytes())
    if os.environ.get("ITCPD_PEER", "1") == "1":
        # fused all-reduce + solve over NVLink peer memory: exchange the CUDA IPC handles of the exchange buffers
        if "ITCPD_PEER_GRAPH" in os.environ:  # default on: NCCL-free sweeps (device-side epochs), CUDA-graph replay
            eng.set_option("peer_graph", int(os.environ["ITCPD_PEER_GRAPH"] != "0"))
        mine = torch.frombuffer(bytearray(eng.peer_export()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        eng.peer_import(world, rank, b"".join(h.cpu().numpy().tobytes() for h in allh))
    eng.compute_grams()
    nT = eng.tensor_norm()
    inner, norm2 = eng.sweep(nsweeps)
    fits = 1.0 - np.sqrt(np.abs(nT * nT + norm2 - 2 * np.abs(inner))) / nT
    last = eng.allgather_factor(len(dims) - 1, dims[-1])
    f0 = eng.get_factor(0)
    ok = True
    if rank == 0:
        ref = itcpd.Engine(local)
        ref.generate_tensor(dims, seed=7)
        ref.set_cpd(factors, np.ones(R))
        ref.compute_grams()
        nT1 = ref.tensor_norm()
        i1, n1 = ref.sweep(nsweeps)
        fits1 = 1.0 - np.sqrt(np.abs(nT1 * nT1 + n1 - 2 * np.abs(i1))) / nT1
        d_fit = float(np.max(np.abs(fits - fits1)))
        d_last = float(np.linalg.norm(last - ref.get_factor(len(dims) - 1)) / np.linalg.norm(last))
        d_f0 = float(np.linalg.norm(f0 - ref.get_factor(0)) / np.linalg.norm(f0))
        print(f"world={world} |norm diff|={abs(nT - nT1):.3e} max|dfit|={d_fit:.3e} last-factor rel diff={d_last:.3e} factor0 rel diff={d_f0:.3e}")
        ok = abs(nT - nT1) < 1e-10 * nT1 and d_fit < 1e-10 and d_last < 1e-8 and d_f0 < 1e-8
        ref.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("MULTI_GPU_OK")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
