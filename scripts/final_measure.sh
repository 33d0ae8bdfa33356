#!/bin/bash
# Round-end measurement pass on ONE GPU (run under gpurun): bench lines of both arms, ncu launch lists, ncu --set full captures of the
# dominant kernels, the configs report.  Everything lands in gpurun_out/ with the prefix given as $1.
cd "$(dirname "$0")/.."
P=${1:-r2c}; O=gpurun_out; mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_reference.json 2> $O/${P}_ref.err
python bench.py > $O/${P}_bench_1gpu.json 2> $O/${P}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > $O/${P}_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${P}_cycle_launches.csv python scripts/cycle_profile.py > $O/${P}_ncu_cycle.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:admm_kernel_tm -s 3 -c 1 -f -o $O/${P}_prof_n30 python scripts/tune.py > $O/${P}_ncu_n30.log 2>&1
TUNE_N=50 ncu --set full --clock-control none --import-source on -k regex:admm_kernel_tmw -s 3 -c 1 -f -o $O/${P}_prof_n50 python scripts/tune.py > $O/${P}_ncu_n50.log 2>&1
TUNE_N=100 ncu --set full --clock-control none --import-source on -k regex:admm_kernel_tmw -s 3 -c 1 -f -o $O/${P}_prof_n100 python scripts/tune.py > $O/${P}_ncu_n100.log 2>&1
TUNE_N=50 TUNE_RATE=0.032 ncu --set full --clock-control none --import-source on -k regex:admm_kernel_tmw -s 3 -c 1 -f -o $O/${P}_prof_rate50 python scripts/tune.py > $O/${P}_ncu_rate50.log 2>&1
python bench.py --configs --configs-out $O/${P}_configs_report.json > $O/${P}_configs.log 2>&1
ls -la $O | tail -20
