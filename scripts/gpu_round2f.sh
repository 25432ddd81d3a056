#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "relin or circuit_b" 2>&1 | tail -3
echo "== bulk"; timeout 300 python scripts/square_relin_probe.py --nq 2048
echo "== ldg"; PPLP_RELIN_BULK=0 timeout 300 python scripts/square_relin_probe.py --nq 2048
