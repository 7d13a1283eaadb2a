#!/bin/bash
# Build libb200ode.so (sm_100a only) and the hardware probe in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
$NVCC $FLAGS -shared -o ../libb200ode.so b200ode.cu -cudart static 2> build_ptxas.log || { cat build_ptxas.log; exit 1; }
# hardware probes (operand layouts, MMA rates, TMA box layout): results under profiles/
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o umma_probe umma_probe.cu
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o umma_rate umma_rate.cu
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o tma_layout_probe tma_layout_probe.cu -lcuda
echo "built $(ls -la ../libb200ode.so)"
