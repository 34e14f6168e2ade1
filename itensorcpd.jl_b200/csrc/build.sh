#!/usr/bin/env bash
# Builds libitcpd_b200.so for sm_100a (the analogue of the reference's deps/build.jl:15-17, with nvcc).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}" "${HERE}/_obj"
NVCC="${NVCC:-nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
pids=()
for f in api gemm_dmma gemm_i8 kernels solve qrcp qrcp_wide sampled sampled_sharded peer_graph comm sparse_sign; do
  src="${HERE}/${f}.cu"; obj="${HERE}/_obj/${f}.o"
  stale=0
  for dep in "${src}" "${HERE}"/*.cuh "${HERE}/../../include/itcpd_b200.h"; do
    [[ ! -f "${obj}" || "${dep}" -nt "${obj}" ]] && stale=1
  done
  if [[ "${stale}" == 1 ]]; then
    "${NVCC}" "${FLAGS[@]}" -c "${src}" -o "${obj}" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -shared -o "${OUT}/libitcpd_b200.so" "${HERE}"/_obj/*.o -cudart static -ldl -lpthread -lrt
echo "built ${OUT}/libitcpd_b200.so"
