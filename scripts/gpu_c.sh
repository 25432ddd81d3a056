#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "circuit_b or multiply or square or behzf or fp64_base or primitives" 2>&1 | tail -3
timeout 300 python scripts/circuit_b_probe.py --chunk 256
timeout 300 python scripts/circuit_b_probe.py --chunk 512
timeout 300 python scripts/square_only_probe.py --nq 2048 --reps 10
