#!/bin/bash
# A/B timing of tuning builds (f110-mpc_b200/tune_libs/*.so, git-ignored) against the in-tree library: scripts/tune.py per library and case.
cd "$(dirname "$0")/.."
out=gpurun_out/ab_tune.txt; mkdir -p gpurun_out; : > $out
run() { echo "## $*" >> $out; env "$@" python scripts/tune.py 2>&1 | tail -1 >> $out; }
cur=$PWD/f110-mpc_b200/libf110mpc_b200.so
for n in 30 40 50 63 100; do run F110_LIB=$cur TUNE_N=$n; done
run F110_LIB=$cur TUNE_N=50 TUNE_RATE=0.032
run F110_LIB=$cur TUNE_N=30 TUNE_RATE=0.032
cat $out
