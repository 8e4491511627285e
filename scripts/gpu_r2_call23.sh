#!/bin/bash
# CTA-pair (cta_group::2) GEMM: exactness, then timing against the single-CTA persistent kernel
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gemm_gpu.py -q -m gpu --timeout 100 -k tcgen05 -s 2>&1 | tail -14 | cut -c1-250 | tee gpurun_out/pytest_pair.log
timeout -k 5 100 python scripts/pair_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/pair_bench.log
