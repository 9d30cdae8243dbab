#!/usr/bin/env bash
# Builds the C-ABI shared library for sm_100a, in-tree:  meshflow_b200/libmeshflow_b200.so
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libmeshflow_b200.so"
NVCC="${NVCC:-nvcc}"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a
       -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=hidden --expt-relaxed-constexpr ${MF_NVCC_EXTRA:-})
objs=()
pids=()
for f in cabi vertex_motion jacobi warp stability; do
  rm -f "${here}/${f}.o"
  "${NVCC}" "${FLAGS[@]}" -c "${here}/${f}.cu" -o "${here}/${f}.o" &
  pids+=($!)
  objs+=("${here}/${f}.o")
done
for p in "${pids[@]}"; do wait "$p" || { echo "nvcc failed" >&2; exit 1; }; done
"${NVCC}" -shared -gencode arch=compute_100a,code=sm_100a -o "${out}" "${objs[@]}"
echo "built ${out}"
