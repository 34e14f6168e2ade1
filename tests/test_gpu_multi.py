"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): slab-sharded ALS + NCCL == single GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_gpu_slab_sharding_matches_single_gpu(peer):
    """peer=1: the M_n all-reduce is fused into the row-solve kernel over NVLink peer memory (CUDA IPC), host-counted exchange epochs,
    kernel-by-kernel launches (peer_graph off); peer=0: plain NCCL all-reduce.  Both must reproduce the single-GPU trajectory."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = ["timeout", "120", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "2961" + ("7" if peer == "1" else "8"), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, ITCPD_PEER=peer, ITCPD_PEER_GRAPH="0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_gpu_sharded_sampled_path_matches_single_gpu():
    """csrc/sampled_sharded.cu: owner-rank gathers + all-reduced sampled normal equations == the single-GPU sampled update."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = ["timeout", "120", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29619", os.path.join(ROOT, "tools", "multi_gpu_sampled_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MULTI_GPU_SAMPLED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_gpu_peer_graph_matches_single_gpu():
    """csrc/peer_graph.cu (the default sharded path): device-side exchange epochs + peer small all-reduce, the sharded sweep replayed
    from a CUDA graph."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = ["timeout", "120", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29620", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, ITCPD_PEER="1", ITCPD_PEER_GRAPH="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
