#!/bin/bash
# One GPU-box pass: parity tests, smoke, the default bench, then ncu (launch list + full capture of the headline kernel).
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 3 --queries 1024 --e2e-queries 64 --no-extras --cpu-sample 8"
timeout 600 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
timeout 600 $SMALL > $OUT/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:circuit_a_kernel -s 3 -c 2 -f -o $OUT/prof_circuit_a_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -20
