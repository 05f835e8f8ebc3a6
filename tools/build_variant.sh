#!/bin/bash
# build exoplanet_b200/libxo_b200_<name>.so with one source recompiled under extra -D flags (tuning A/B on the GPU box:
# XO_B200_LIB=exoplanet_b200/libxo_b200_<name>.so python bench.py ...)
# usage: tools/build_variant.sh name source.cu -DFLAG=1 [...]
set -e
NAME=$1; SRC=$2; shift 2
cd "$(dirname "$0")/../exoplanet_b200/csrc"
python build.py > /dev/null
OBJS=""
for f in xo_ops xo_transit xo_model xo_celerite xo_host xo_diag xo_joint xo_sampler xo_nuts; do
  if [ "$f.cu" = "$SRC" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin /usr/bin/g++ "$@" -c $f.cu -o /tmp/${f}_$NAME.o
    OBJS="$OBJS /tmp/${f}_$NAME.o"
  else
    OBJS="$OBJS _build/$f.o"
  fi
done
nvcc -shared -o ../libxo_b200_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -cudart shared -ldl
echo ../libxo_b200_$NAME.so
