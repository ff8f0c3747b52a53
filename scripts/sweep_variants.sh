#!/bin/bash
# time a set of classes with each library variant (LB200_LIB_SUFFIX)
for suf in "" _mb3 _mb4; do
  for c in "2 2 2 2" "2 1 2 2" "2 1 2 1" "1 1 2 2" "2 0 2 2" "1 1 1 1" "0 0 2 2" "1 0 1 0" "0 0 0 0"; do
    echo -n "variant '$suf' "; LB200_LIB_SUFFIX=$suf python scripts/prof_class.py $c 1048576 3 | tail -1
  done
done
