#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_r02f.log
timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
timeout 300 python scripts/circuit_b_probe.py --chunk 256
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
