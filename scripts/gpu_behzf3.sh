#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
K="multiply or square or circuit_b or relin or batch_encoder"
timeout 900 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -3
timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active --clock-control none -c 12 --csv --log-file $OUT/r02_squaref_launches.csv python scripts/square_only_probe.py --nq 512 --reps 1 > $OUT/ncu_squaref.log 2>&1
timeout 900 python bench.py 2> $OUT/bench_r02d.err > $OUT/bench_r02d.json; tail -3 $OUT/bench_r02d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02d.json'))
print(d['value'], d['e2e']['value'], json.dumps(d['extras']['circuit_b'])[:600])
print(json.dumps(d['extras']['config4_sweep']))
PY
