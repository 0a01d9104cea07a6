#!/bin/bash
# build_variant.sh NAME [-DMACRO ...]: a build of the library with extra macros into gpurun_ab/NAME.so (travels with gpurun)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p gpurun_ab build_ab/$name
rm -f build_ab/$name/*.o
pids=
for f in capi multi exposure_grad lp_grad lp_grad_mom ppc sampler_kernels advi nuts nuts_batched; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
      -c -o build_ab/$name/$f.o ppcseq_b200/csrc/$f.cu &
  pids="$pids $!"
done
for p in $pids; do wait $p || { echo "compile failed"; exit 1; }; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_ab/$name.so build_ab/$name/*.o -lcudart -lpthread
echo built gpurun_ab/$name.so
