#!/bin/bash
# time every kernel build variant under circuitree_b200/variants on 2^20 repressilator trajectories
for lib in circuitree_b200/variants/*.so; do
  echo "== $lib"
  CIRCUITREE_B200_LIB=$PWD/$lib python scripts/prof_run.py ${1:-20} 2
done
