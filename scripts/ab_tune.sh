#!/bin/bash
# A/B timing of tuning builds (f110-mpc_b200/tune_libs/*.so, git-ignored) against the in-tree library: scripts/tune.py per library and case.
cd "$(dirname "$0")/.."
out=gpurun_out/ab_tune.txt; mkdir -p gpurun_out; : > $out
run() { echo "## $*" >> $out; env "$@" python scripts/tune.py 2>&1 | tail -1 >> $out; }
cur=$PWD/f110-mpc_b200/libf110mpc_b200.so
for n in 32 40 50 63; do
  run F110_LIB=$PWD/f110-mpc_b200/tune_libs/base.so TUNE_N=$n TUNE_RATE=0.032
  run F110_LIB=$cur TUNE_N=$n TUNE_RATE=0.032
  run F110_LIB=$cur TUNE_N=$n TUNE_RATE=0.032
done
run F110_LIB=$cur TUNE_N=50
run F110_LIB=$cur TUNE_N=63
cat $out
