#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "encrypt or protocol or proximity or keygen or decrypt or shim" 2>&1 | tail -4
echo "== new"; timeout 300 python scripts/protocol_probe.py --nq 4096 --reps 5 2>&1 | tail -2
echo "== old inverse"; PPLP_ENC_INV32=0 timeout 300 python scripts/protocol_probe.py --nq 4096 --reps 5 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 80 --csv --log-file $OUT/r02_protocol_launches.csv python scripts/protocol_probe.py --nq 2048 --reps 1 > $OUT/ncu_proto.log 2>&1
tail -1 $OUT/ncu_proto.log
