#!/usr/bin/env bash
# Builds the C-ABI shared library for sm_100a, in-tree:  meshflow_b200/libmeshflow_b200.so
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libmeshflow_b200.so"
NVCC="${NVCC:-nvcc}"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a
       -Xcompiler -fPIC,-O2,-ffp-contract=off --expt-relaxed-constexpr ${MF_NVCC_EXTRA:-})
objs=()
for f in cabi vertex_motion jacobi warp stability; do
  "${NVCC}" "${FLAGS[@]}" -c "${here}/${f}.cu" -o "${here}/${f}.o" &
  objs+=("${here}/${f}.o")
done
wait
"${NVCC}" -shared -gencode arch=compute_100a,code=sm_100a -o "${out}" "${objs[@]}"
echo "built ${out}"
