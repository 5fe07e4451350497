#!/bin/bash
# TEST INFRASTRUCTURE: builds tests/_build/libopenblas_b200_hostsim.so -- the library's unchanged host
# code (interface_*.c, runtime.cu compiled by g++ with -DB200_HOSTSIM) on top of a host-memory stand-in
# for the CUDA runtime and CPU stand-ins for the kernels (the oracle).  No nvcc, no GPU, no libcudart.
# cuda_shim.cpp is compiled with protected visibility: the stand-in cuda* functions must win inside this
# library even when a real libcudart is already loaded in the process (torch), while xerbla_ stays
# overridable by the program, as in the product.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$ROOT/tests/_build${HOSTSIM_SANITIZE:+/$HOSTSIM_SANITIZE}"
SAN="${HOSTSIM_SANITIZE:+-fsanitize=$HOSTSIM_SANITIZE -fno-omit-frame-pointer}"      # HOSTSIM_SANITIZE=address|thread: instrumented build
mkdir -p "$OUT"
make -s -C "$ROOT/oracle" _build/liboracle.so
CSRC="$ROOT/openblas_b200/csrc"
INC="-I/usr/local/cuda/include -I$CSRC -I$ROOT/include"
g++ $SAN -std=c++17 -O1 -g -fPIC -fvisibility=hidden -DB200_HOSTSIM $INC -x c++ -c "$CSRC/runtime.cu" -o "$OUT/runtime.o"
g++ $SAN -std=c++17 -O1 -g -fPIC -fvisibility=hidden -DB200_HOSTSIM $INC -x c++ -c "$CSRC/summa.cu" -o "$OUT/summa.o"
g++ $SAN -std=c++17 -O1 -g -fPIC -fvisibility=hidden -DB200_HOSTSIM $INC -c "$ROOT/tests/hostsim/sim_kernels.cpp" -o "$OUT/sim_kernels.o"
g++ $SAN -std=c++17 -O1 -g -fPIC -fvisibility=protected -DB200_HOSTSIM $INC -c "$ROOT/tests/hostsim/cuda_shim.cpp" -o "$OUT/cuda_shim.o"
for f in interface_gemm interface_level3 xerbla control; do
  gcc $SAN -O1 -g -fPIC -fvisibility=hidden -std=gnu11 $INC -c "$CSRC/$f.c" -o "$OUT/$f.o"
done
g++ $SAN -shared -o "$OUT/libopenblas_b200_hostsim.so" "$OUT"/runtime.o "$OUT"/summa.o "$OUT"/sim_kernels.o "$OUT"/cuda_shim.o "$OUT"/interface_gemm.o \
    "$OUT"/interface_level3.o "$OUT"/xerbla.o "$OUT"/control.o -L"$ROOT/oracle/_build" -loracle -Wl,-rpath,"$ROOT/oracle/_build" -lpthread -ldl
echo "$OUT/libopenblas_b200_hostsim.so"
