#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
K="multiply or square or circuit_b or relin or batch_encoder"
echo "== default (FP64 base, fused)"; timeout 900 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -4
for nq in 512 2048; do
timeout 300 python scripts/square_only_probe.py --nq $nq --reps 10
done
./build/butterfly_f64_mb 2>&1 | grep -E "threads  256|threads  512"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active --clock-control none -c 12 --csv --log-file $OUT/r02_squaref_launches.csv python scripts/square_only_probe.py --nq 512 --reps 1 > $OUT/ncu_squaref.log 2>&1
grep -E "gpu__time|fp64|per_cycle" $OUT/r02_squaref_launches.csv | tail -12 | cut -d, -f5,13,15 | cut -c1-200
