#!/usr/bin/env bash
# build_variant.sh <name> <extra nvcc flags...>  -> gpurun_out/variants/lib_<name>.so
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/../meshflow_b200/csrc" && pwd)"
name=$1; shift
out="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)/variants"; mkdir -p "$out"
tmp=$(mktemp -d)
pids=()
for f in cabi vertex_motion jacobi warp stability; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2,-ffp-contract=off --expt-relaxed-constexpr "$@" -c "$here/$f.cu" -o "$tmp/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p" || { echo "nvcc failed" >&2; exit 1; }; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out/lib_${name}.so" "$tmp"/*.o
rm -rf "$tmp"; echo "built $out/lib_${name}.so"
