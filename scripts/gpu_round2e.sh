#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "encrypt or protocol or proximity" 2>&1 | tail -3
timeout 300 python scripts/protocol_probe.py --nq 4096 --reps 5 2>&1 | tail -1
timeout 900 python bench.py 2> $OUT/bench_r02e.err > $OUT/bench_r02e.json; tail -2 $OUT/bench_r02e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02e.json'))
x=d['extras']
print(d['value'], d['roofline']['frac'], d['e2e']['value'], x['protocol_e2e']['value'], x['circuit_b']['groups_per_s'], x['ntt_gbs'], x['intt_gbs'])
PY
timeout 600 python scripts/square_only_probe.py --nq 512 --reps 1 > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:behzf -s 8 -c 4 -f -o $OUT/r02_behzf python scripts/square_only_probe.py --nq 512 --reps 1 > $OUT/ncu_behzf_full.log 2>&1
timeout 600 python scripts/protocol_probe.py --nq 2048 --reps 1 > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:enc32_inverse -s 2 -c 2 -f -o $OUT/r02_enc32inv python scripts/protocol_probe.py --nq 2048 --reps 1 > $OUT/ncu_enc32_full.log 2>&1
ls -la $OUT/*.ncu-rep | tail -4
