#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $OUT/pytest_r02d.log
for c in 256 512; do timeout 300 python scripts/circuit_b_probe.py --chunk $c; done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 60 --csv --log-file $OUT/r02_circuit_b_launches.csv python scripts/circuit_b_probe.py --reps 1 > $OUT/ncu_cb.log 2>&1
tail -1 $OUT/ncu_cb.log
