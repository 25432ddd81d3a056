#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_r02d.log
echo "== default build"; timeout 300 python scripts/ntt_bench.py --n 8192 16384 2>&1 | tail -3; timeout 300 python scripts/square_relin_probe.py --nq 2048
echo "== FRND build"; export PPLP_B200_LIB=$PWD/build/frnd/libpplp_b200_frnd.so
timeout 300 python scripts/ntt_bench.py --n 8192 16384 2>&1 | tail -3; timeout 300 python scripts/square_relin_probe.py --nq 2048
timeout 900 python -m pytest tests -m gpu -q -x -k "ntt or multiply or square or relin or encrypt or decrypt" 2>&1 | tail -3
