#!/bin/bash
# usage: tools/k2_ab.sh build|run   — variants: name:flags
set -e
cd "$(dirname "$0")"
VARIANTS=${VARIANTS:-"base: oldexp:-DSO_EXP_OLD"}
if [ "$1" = build ]; then
  mkdir -p ab
  for v in $VARIANTS; do
    name=${v%%:*}; flags=${v#*:}; flags=${flags//,/ }
    ( nvcc -std=c++20 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 $flags -o ab/k2_$name k2_ab.cu && echo built $name ) &
  done
  wait
else
  for rep in 1 2; do for v in $VARIANTS; do name=${v%%:*}; ./ab/k2_$name ${ROWS:-1000000000} $name ${KIND:-2} ${RANGE:-0} ${LAUNCHES:-23} ${NALPHA:-1}; done; done
fi
