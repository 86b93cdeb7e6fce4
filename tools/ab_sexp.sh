#!/bin/bash
# A/B of the scaled-product exponential (so_exps) against the previous build: tools/ab/k2_base (HEAD~) vs tools/ab/k2_sexp
cd "$(dirname "$0")"
for rep in 1 2; do
  for spec in "2 1 1" "2 0 1" "1 1 1" "4 1 1" "5 1 1" "3 1 1" "3 1 2" "3 1 4" "6 0 1"; do
    set -- $spec
    for v in base sexp; do ./ab/k2_$v 1000000000 "$v-kind$1-range$2-k$3" $1 $2 23 $3; done
  done
done
