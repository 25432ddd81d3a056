#!/bin/bash
# GPU pass for the FP64-base BEHZ product: parity in the three modes, throughput probe, per-kernel launch list.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
K="multiply or square or circuit_b or relin or batch_encoder"
echo "== default (FP64 base, fused)"; timeout 900 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -8
echo "== unfused"; PPLP_BEHZF_FUSED=0 timeout 900 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -8
echo "== 61-bit base"; PPLP_BEHZ_BASE=61 timeout 900 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -8
for nq in 512 2048; do
echo "== probe nq=$nq: f64 fused / unfused / 61"
timeout 300 python scripts/square_only_probe.py --nq $nq --reps 10
PPLP_BEHZF_FUSED=0 timeout 300 python scripts/square_only_probe.py --nq $nq --reps 10
PPLP_BEHZ_BASE=61 timeout 300 python scripts/square_only_probe.py --nq $nq --reps 10
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 60 --csv --log-file $OUT/r02_squaref_launches.csv python scripts/square_only_probe.py --nq 512 --reps 1 > $OUT/ncu_squaref.log 2>&1
tail -12 $OUT/r02_squaref_launches.csv | cut -c1-400
