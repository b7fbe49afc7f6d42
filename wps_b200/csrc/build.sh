#!/bin/bash
# Builds wps_b200/libwpsgpu.so (sm_100a).  IEEE-strict device flags: the reference arithmetic is
# binary32 without FMA contraction or flush-to-zero (hard part 4).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="$EXTRA_NVCC_FLAGS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ftz=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-ffp-contract=off,-O2"
mkdir -p build
pids=()
for f in api metgrid geogrid misc mpas; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer build/$f.o)" ] || [ ../../include/wpsgpu.h -nt build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libwpsgpu.so build/api.o build/metgrid.o build/geogrid.o build/misc.o build/mpas.o -lcudart
echo "built $(cd .. && pwd)/libwpsgpu.so"
# C++ host side above the C ABI (include/wpshost.h): links against libwpsgpu.so next to it
g++ -O2 -std=c++17 -fPIC -shared -Wall -Wextra -ffp-contract=off -o ../libwpshost.so host/wpshost.cpp -L.. -lwpsgpu -Wl,-rpath,'$ORIGIN'
echo "built $(cd .. && pwd)/libwpshost.so"
