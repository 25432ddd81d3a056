#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "ntt or wide_fp64" 2>&1 | tail -3
echo "== cluster"; timeout 300 python scripts/ntt_bench.py --n 16384 2>&1 | tail -1
echo "== single CTA"; PPLP_NTT_CLUSTER=0 timeout 300 python scripts/ntt_bench.py --n 16384 2>&1 | tail -1
