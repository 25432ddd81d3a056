#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_r02e.log
timeout 300 python scripts/circuit_b_probe.py --chunk 256
timeout 300 python scripts/circuit_b_probe.py --chunk 512
