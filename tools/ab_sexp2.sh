#!/bin/bash
# checked (range-check) kernels: previous build vs scaled-product exp with the range check on the factors
cd "$(dirname "$0")"
for rep in 1 2; do
  for spec in "2 0 1" "2 1 1" "1 0 1" "4 0 1" "5 0 1" "3 0 2"; do
    set -- $spec
    for v in base sexp2; do ./ab/k2_$v 1000000000 "$v-kind$1-range$2-k$3" $1 $2 23 $3; done
  done
done
