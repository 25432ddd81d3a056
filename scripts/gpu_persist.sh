#!/bin/bash
OUT=gpurun_out
echo "== persistent forward"; PPLP_BEHZF_PERSIST=1 timeout 600 python -m pytest tests -m gpu -q -x -k "multiply or square or behzf or fp64_base" 2>&1 | tail -3
PPLP_BEHZF_PERSIST=1 timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
echo "== default"; timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
PPLP_BEHZF_PERSIST=1 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 12 --csv --log-file $OUT/persist_launches.csv python scripts/square_only_probe.py --nq 512 --reps 1 > /dev/null 2>&1
grep -E "forward" $OUT/persist_launches.csv | tail -2 | cut -d, -f5,13-15
