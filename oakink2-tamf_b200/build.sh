#!/usr/bin/env bash
# Builds libtamf_b200.so in-tree for sm_100a.  Usage: oakink2-tamf_b200/build.sh [-v]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$here/lib" "$here/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       -Xcudafe --diag_suppress=177 --expt-relaxed-constexpr)
if [[ "${1:-}" == "-v" ]]; then FLAGS+=(-Xptxas -v); fi
pids=()
for f in api encoder nn mano denoiser refiner; do
  src="$here/csrc/$f.cu"; obj="$here/build/$f.o"
  newest=$(ls -t "$here"/csrc/*.cu "$here"/csrc/*.cuh "$here"/../include/*.h | head -1)
  if [[ ! -f "$obj" || "$newest" -nt "$obj" || "${1:-}" == "-v" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$here/lib/libtamf_b200.so" "$here"/build/{api,encoder,nn,mano,denoiser,refiner}.o -lcudart_static -lpthread -ldl -lrt
echo "built $here/lib/libtamf_b200.so"
