#!/bin/bash
# Build A/B variants of the library for kernel experiments: tools/build_variants.sh name "<extra nvcc flags>" ...
# Each variant lands in hopfield_network_go_b200/lib/variants/<name>.so (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p hopfield_network_go_b200/lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -diag-suppress 177 $flags \
    -o hopfield_network_go_b200/lib/variants/$name.so hopfield_network_go_b200/csrc/api.cu hopfield_network_go_b200/csrc/host_rng.cpp -ldl &
done
wait
ls -la hopfield_network_go_b200/lib/variants/
