#!/bin/bash
# round 2, multi-GPU parity tests with the final code (run with gpurun --gpus 2 or 8): NCCL decomposed == single block, transient loops over
# ranks, the unmodified MPI program test/heavisidefilter.cpp (filter ghost exchange) — whatever the box has GPUs for
mkdir -p gpurun_out
N=$(python -c "from panslbm2_b200 import _lib; print(_lib.lib().pl_device_count())")
(timeout 1200 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_transient.py tests/test_gpu_dropin.py -m gpu -q -k "nccl or over_ranks or heavisidefilter" --maxfail=10 --timeout=500 -rs 2>&1 | tail -30) > gpurun_out/r02k_tests_${N}gpus.log
cat gpurun_out/r02k_tests_${N}gpus.log
