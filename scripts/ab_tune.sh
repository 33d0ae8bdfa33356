#!/bin/bash
# A/B timing of tuning builds (f110-mpc_b200/tune_libs/*.so, git-ignored) against the in-tree library: scripts/tune.py per library and case.
cd "$(dirname "$0")/.."
out=gpurun_out/ab_tune.txt; mkdir -p gpurun_out; : > $out
run() { echo "## $*" >> $out; env "$@" python scripts/tune.py 2>&1 | tail -1 >> $out; }
cur=$PWD/f110-mpc_b200/libf110mpc_b200.so
for lib in $PWD/f110-mpc_b200/tune_libs/fac2.so $cur $PWD/f110-mpc_b200/tune_libs/fac2.so $cur; do
  run F110_LIB=$lib
done
for lib in $PWD/f110-mpc_b200/tune_libs/fac2.so $cur; do
  run F110_LIB=$lib TUNE_B=1024
  run F110_LIB=$lib TUNE_N=50
  run F110_LIB=$lib TUNE_N=100
done
cat $out
