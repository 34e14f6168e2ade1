"""Host-side dry run of the whole C-ABI without a GPU.

The product objects (csrc/_obj/*.o, exactly what libitcpd_b200.so is linked from) are linked a second time with
`-cudart shared`; the process that loads that copy finds tests/fake_cudart.cpp under the names libcudart.so.12 and
libnccl.so.2.  The fake runtime never executes a kernel: it checks what makes real launches, copies, TMA descriptor encodes and
stream captures fail (see its header).  tests/dry_run_driver.py then drives the real entry points over the edge-case shapes,
every runtime option, BASELINE.json's configs A-D at FULL size (lazily mapped memory), their per-rank slabs, two ranks in one
process (NCCL-only, fused peer solve, peer_graph, sharded sampled path) and an out-of-memory device.
This is how the host logic of the paths written after the round's GPU budget was spent (gemm_i8 incl. split-K and the
does-not-fit fallbacks, early_pass_b, peer_graph, the sharded sampled path) was executed at all; it found two defects on its
first runs (a grid.y overflow in the INT8 exponent pre-pass for tiny leading modes, and a stale "last error" after a failed
allocation that made the next launch check fail).  It says nothing about numerics or speed."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "_obj")
DRY = os.path.join(ROOT, "oracle", "_build", "dry")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None or not os.path.isdir(CUDA_INC),
                                reason="needs the CUDA toolkit (nvcc link + cuobjdump resource table); no GPU")


@pytest.fixture(scope="module")
def dry_results():
    subprocess.run(["bash", os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "build.sh")], check=True, capture_output=True)   # up-to-date objects
    os.makedirs(DRY, exist_ok=True)
    objs = sorted(os.path.join(OBJ, f) for f in os.listdir(OBJ) if f.endswith(".o"))
    lib = os.path.join(DRY, "libitcpd_dry.so")
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-cudart", "shared", "-ldl", "-lpthread", "-lrt"], check=True, capture_output=True)
    # kernel resource table (registers, static shared memory) for the launch checks
    out = subprocess.run(["cuobjdump", "-res-usage", lib], check=True, capture_output=True, text=True).stdout.splitlines()
    rows, name = [], None
    for line in out:
        line = line.strip()
        if line.startswith("Function "):
            name = line[len("Function "):].rstrip(":")
        elif name and "REG:" in line:
            kv = dict(p.split(":") for p in line.split() if ":" in p)
            rows.append(f"{name} {kv.get('REG', 0)} {kv.get('SHARED', 0)}")
            name = None
    assert len(rows) >= 90, len(rows)
    table = os.path.join(DRY, "kernels.tbl")
    open(table, "w").write("\n".join(rows) + "\n")
    vermap = os.path.join(DRY, "ver.map")
    open(vermap, "w").write("libcudart.so.12 { global: *; };\n")
    fake = os.path.join(DRY, "libcudart.so.12")
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "fake_cudart.cpp"), "-I", CUDA_INC,
                    f"-Wl,--version-script={vermap}", "-Wl,-soname,libcudart.so.12", "-o", fake], check=True, capture_output=True)
    link = os.path.join(DRY, "libnccl.so.2")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink("libcudart.so.12", link)
    env = dict(os.environ, LD_LIBRARY_PATH=DRY + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), FAKECUDA_KERNEL_TABLE=table)
    for k in ("ITCPD_GEMM_I8", "ITCPD_EARLY_B", "ITCPD_CHOL", "ITCPD_NO_GRAPH", "ITCPD_NO_SWIZZLE"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dry_run_driver.py")], env=env, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("DRYRUN_JSON ")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    return json.loads(lines[-1][len("DRYRUN_JSON "):])


SCENARIOS = [
    "small_shapes_default_options", "small_shapes_every_option", "config_A_200cubed_rank50", "config_B_1024cubed_rank64", "config_B_early_pass_b",
    "config_B_gemm_i8_on_the_fly", "config_B_gemm_i8_prepacked_early_pass_b", "config_C_256pow4_rank32", "config_C_gemm_i8_prepacked",
    "config_D_2048cubed_rank128", "config_D_gemm_i8_prepacked_planes_of_one_unfolding_fit_the_other_converts_on_the_fly", "slab_B8_gemm_i8_split_k",
    "slab_D8_gemm_i8_split_k", "two_ranks_nccl_only", "two_ranks_fused_peer_solve", "two_ranks_peer_graph_is_captured_without_nccl",
    "two_ranks_order4_small_with_sampled_path", "two_ranks_gemm_i8", "eight_ranks_config_B_slabs_fused_peer_solve",
    "eight_ranks_config_D_slabs_peer_graph_gemm_i8", "pivot_setup_and_bench_entry_points", "sampled_path_single_rank", "end_to_end_call_from_host_buffers",
    "out_of_memory_is_an_error_code_not_a_crash", "fuzz_dense_shapes_and_options", "fuzz_two_rank_sharded_dense_and_sampled",
    "single_sweep_calls_replay_a_graph_only_with_the_option",
    "chol_alg_3_uses_the_right_looking_kernel_only_where_the_factorisation_is_exposed",
]


def test_driver_ran_every_scenario(dry_results):
    assert sorted(dry_results) == sorted(SCENARIOS)


@pytest.mark.parametrize("name", SCENARIOS)
def test_dry_run_scenario(dry_results, name):
    r = dry_results[name]
    assert r["ok"], r["error"]
    assert r["violations"] == [], r["violations"][:10]
    assert r["leaked_device_allocs"] == 0
    assert r["launches"] > 0


def test_fake_runtime_reports_what_it_is_meant_to_catch(dry_results):
    """negative control: the validating runtime itself flags an illegal launch, an overrunning copy and an unjoined capture"""
    src = r'''
#include <cuda_runtime_api.h>
#include <cstdio>
extern "C" int fakecuda_violation_count();
extern "C" void **__cudaRegisterFatBinary(void *);
extern "C" void __cudaRegisterFunction(void **, const char *, char *, const char *, int, uint3 *, uint3 *, dim3 *, dim3 *, int *);
static char fn;
int main() {
    __cudaRegisterFunction(__cudaRegisterFatBinary(nullptr), &fn, nullptr, "control_kernel", 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    int before = fakecuda_violation_count();
    cudaLaunchKernel(&fn, dim3(1, 70000, 1), dim3(256, 1, 1), nullptr, 0, nullptr);          // grid.y too large
    cudaLaunchKernel(&fn, dim3(1, 1, 1), dim3(256, 1, 1), nullptr, 100 * 1024, nullptr);      // dynamic smem without the attribute
    void *d; cudaMalloc(&d, 1000); char h[2000];
    cudaMemcpy(d, h, 2000, cudaMemcpyHostToDevice);                                           // overruns the allocation
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, 0); cudaStreamCreateWithFlags(&b, 0);
    cudaEvent_t e; cudaEventCreate(&e);
    cudaGraph_t g;
    cudaStreamBeginCapture(a, cudaStreamCaptureModeRelaxed);
    cudaLaunchKernel(&fn, dim3(1, 1, 1), dim3(32, 1, 1), nullptr, 0, a);
    cudaEventRecord(e, a); cudaStreamWaitEvent(b, e, 0);
    cudaLaunchKernel(&fn, dim3(1, 1, 1), dim3(32, 1, 1), nullptr, 0, b);                       // forked, never joined
    cudaError_t r = cudaStreamEndCapture(a, &g);
    printf("%d %d\n", fakecuda_violation_count() - before, (int)(r != cudaSuccess));
    return 0;
}
'''
    cpp, exe = os.path.join(DRY, "control.cpp"), os.path.join(DRY, "control")
    open(cpp, "w").write(src)
    subprocess.run(["g++", "-std=c++17", cpp, "-I", CUDA_INC, "-L", DRY, "-l:libcudart.so.12", f"-Wl,-rpath,{DRY}", "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=DRY)).stdout.split()
    assert out == ["4", "1"], out


@pytest.mark.parametrize("env_extra,args", [({}, []), ({"ITCPD_GEMM_I8": "2", "ITCPD_EARLY_B": "1"}, []), ({}, ["--config", "B8"]), ({}, ["--config", "A"])])
def test_bench_line_contract_in_the_dry_run(dry_results, env_extra, args):
    """bench.py's own arm at N = 1, start to finish, behind the fake runtime (numbers are meaningless, the JSON contract is not)"""
    wrapper = (f"import sys, runpy; sys.path.insert(0, {ROOT!r}); import itcpd; "
               f"itcpd.package._lib.LIB_PATH = {os.path.join(DRY, 'libitcpd_dry.so')!r}; "
               f"sys.argv = ['bench.py', '--steps', '4', '--warmup', '3', '--no-cpu'] + {args!r}; "
               f"runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__')")
    env = dict(os.environ, LD_LIBRARY_PATH=DRY, FAKECUDA_KERNEL_TABLE=os.path.join(DRY, "kernels.tbl"), **env_extra)
    p = subprocess.run([sys.executable, "-W", "ignore", "-c", wrapper], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-1500:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "clocks", "gpu_launches", "roofline", "e2e"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] >= 3 and line["unit"] == "sweeps/s"
    assert line["dtype"].startswith("i8 digits") if env_extra.get("ITCPD_GEMM_I8") else line["dtype"] == "f64"   # the arithmetic the path computes in
    assert set(line["config"]) == {"workload", "l2"}   # the keys the reference arm prints too; everything else lives in config_detail
    assert "parity" in line and "config_detail" in line
    if not args:   # the default workload carries the other BASELINE.json configurations as compact records
        assert [r["config"] for r in line["extra"]] == ["D", "C", "A", "E", "B+gemm_i8=2"], line["extra"]
        for r in line["extra"]:   # (E draws its samples on the device: no kernel runs here, so its host-side pivot check refuses the zeros)
            assert "error" not in r or r["config"] == "E", r
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in line["roofline"], k
    assert line["roofline"]["bound"] == ("hbm" if env_extra.get("ITCPD_GEMM_I8") else "tensor")
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
