#!/bin/bash
# Flappy-Bird family: 3-channel non-square frames + extra features (golden at small size, oracle at real size)
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_learner_gpu.py -q -m gpu --timeout 300 -k "rgb_rect or flappy" -s 2>&1 | tail -30 | cut -c1-400 | tee gpurun_out/pytest_flappy.log
