#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_kernels.log 2>&1
tail -n 6 gpurun_out/pytest_gpu_kernels.log
( time timeout 400 python tools/bench_sustained.py --L 30 --tag r01k ) > gpurun_out/sustained_r01k.log 2>&1
cut -c1-170 gpurun_out/sustained_r01k.log | tail -n 50
