#!/usr/bin/env bash
# Builds libnunet_b200.so (sm_100a only) next to this script.  Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --shared \
    -o libnunet_b200.so engine.cu ${EXTRA_SRCS:-} 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $(pwd)/libnunet_b200.so"
