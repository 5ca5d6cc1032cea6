#!/bin/bash
# Final one-GPU check of round 2 (20 s of GPU time were left): the launchers whose host part was factored out for the
# parameter images (dense DIRECT / TILED / DMMA, folded diagonals, batched diagonals, tile programs) on the B200, through the
# C ABI, against the numpy oracle.  Output: gpurun_out/r02p_gpu_subset.log
mkdir -p gpurun_out
timeout 19 python -u -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider \
  -k "prediag or tile_program or diag_batch or block_structure or many_controls or small_slabs" 2>&1 | tee gpurun_out/r02p_gpu_subset.log | tail -5
