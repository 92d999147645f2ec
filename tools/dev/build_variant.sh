#!/bin/bash
# dev: build a variant of the library with extra nvcc flags for the thread-per-system translation units only
#   tools/dev/build_variant.sh NAME "<extra flags>"   ->  tools/dev/lib/libvar_NAME.so   (run with PHOSKIN_LIB=...)
set -e
cd "$(dirname "$0")/../.."
NAME=$1; shift
EXTRA="$*"
mkdir -p build/var_$NAME tools/dev/lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -Xcompiler -fPIC"
for u in pk_tps_succ pk_tps_dist; do
  nvcc $FLAGS $EXTRA -c -o build/var_$NAME/$u.o phoskintime_b200/csrc/$u.cu 2> build/var_$NAME/$u.log &
done
wait
grep -hE "[1-9][0-9]* bytes (stack frame|spill)" build/var_$NAME/*.log && echo "WARNING: local memory in variant $NAME"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/dev/lib/libvar_$NAME.so build/pk_api.o build/var_$NAME/pk_tps_succ.o build/var_$NAME/pk_tps_dist.o build/pk_dense.o build/pk_global.o build/pk_nlls.o -ldl
echo built tools/dev/lib/libvar_$NAME.so
