#!/bin/bash
# Round-2 GPU pass (one B200): parity in every pipeline variant, throughput probes with their A/B switches, launch lists and full
# ncu captures of the Circuit-B and encryption kernels.  Usage (repo root, under gpurun): bash scripts/gpu_round2.sh
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_r02.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
K="multiply or square or circuit_b or relin or ntt"
for env in "PPLP_BEHZF_FUSED=0" "PPLP_BEHZ_BASE=61" "PPLP_RELIN_BULK=1" "PPLP_BEHZF_PERSIST=1" "PPLP_NTT_CLUSTER=1" "PPLP_ENC_INV32=0"; do
  echo "== parity with $env"; env $env timeout 900 python -m pytest tests -m gpu -q -x -k "$K or encrypt or protocol" 2>&1 | tail -2
done
echo "== probes (default, then each switch)"
timeout 300 python scripts/square_relin_probe.py --nq 2048
PPLP_BEHZ_BASE=61 timeout 300 python scripts/square_relin_probe.py --nq 2048
PPLP_RELIN_BULK=1 timeout 300 python scripts/square_relin_probe.py --nq 2048
PPLP_BEHZF_PERSIST=1 timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
timeout 300 python scripts/circuit_b_probe.py --chunk 256
timeout 300 python scripts/protocol_probe.py --nq 4096 --reps 5 | tail -1
PPLP_ENC_INV32=0 timeout 300 python scripts/protocol_probe.py --nq 4096 --reps 5 | tail -1
timeout 300 python scripts/ntt_bench.py --n 8192 16384 | tail -2
PPLP_NTT_CLUSTER=1 timeout 300 python scripts/ntt_bench.py --n 16384 | tail -1
timeout 900 python bench.py 2> $OUT/bench_r02.err | tee $OUT/bench_r02.json | cut -c1-300
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file $OUT/r02_circuit_b_launches.csv python scripts/circuit_b_probe.py --reps 1 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 80 --csv --log-file $OUT/r02_protocol_launches.csv python scripts/protocol_probe.py --nq 2048 --reps 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:behzf -s 8 -c 4 -f -o $OUT/r02_behzf python scripts/square_only_probe.py --nq 512 --reps 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:enc32_inverse -s 2 -c 2 -f -o $OUT/r02_enc32inv python scripts/protocol_probe.py --nq 2048 --reps 1 > /dev/null 2>&1
ls -la $OUT | tail -12
