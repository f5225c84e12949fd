#!/usr/bin/env bash
# Builds libnunet_b200.so (sm_100a only) and the pybind11 layer _nunet_pybind next to this script.
# Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
PY=${PYTHON:-python}
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --shared \
    -o libnunet_b200.so engine.cu ${EXTRA_SRCS:-} 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $(pwd)/libnunet_b200.so"
# thin pybind11 module over the extern "C" symbols (host code only; finds the library next to itself)
EXT=$("$PY" -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC=$("$PY" -c "import pybind11, sysconfig; print('-I' + pybind11.get_include() + ' -I' + sysconfig.get_paths()['include'])")
g++ -std=c++17 -O2 -shared -fPIC -fvisibility=hidden $INC pybind_module.cpp -o "_nunet_pybind$EXT" \
    -L. -lnunet_b200 -Wl,-rpath,'$ORIGIN'
echo "built $(pwd)/_nunet_pybind$EXT"
