#!/usr/bin/env bash
# Builds libsscg_b200.so (sm_100a only) next to the package. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libsscg_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -shared --expt-relaxed-constexpr -Xptxas -v "$@" \
    -o "${OUT}" "${HERE}/api.cu" -lcudart
echo "built ${OUT}"
