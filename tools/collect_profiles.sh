#!/bin/bash
# GPU box: everything profiles/r02_* is derived from. Usage: bash tools/collect_profiles.sh (under gpurun); outputs in gpurun_out/r2/
set -u
O=gpurun_out/r2
mkdir -p $O
B="python bench.py --steps 8 --warmup 8 --no-reference-gpu --no-cpu-baseline --job 0"
$B > $O/bench_plain.json 2> $O/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv $B > $O/bench_ncu.log 2>&1
python tools/profile_all_once.py > $O/kernels_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_|ss2d_|pointwise_tc3|depthwise3|conv3x3' -o $O/kernels python tools/profile_all_once.py > $O/kernels_ncu.log 2>&1
python tools/run_ss2d_once.py > $O/ss2d_levels.log 2>&1
python tools/run_scan_once.py L0 --bwd > $O/scan_L0.log 2>&1; python tools/run_scan_once.py L1 --bwd >> $O/scan_L0.log 2>&1; python tools/run_scan_once.py L2 --bwd >> $O/scan_L0.log 2>&1
python tools/run_scan_once.py HD --bf16 >> $O/scan_L0.log 2>&1
python tools/run_sbatch.py > $O/sbatch.log 2>&1
ls -la $O
