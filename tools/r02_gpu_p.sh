#!/bin/bash
# Round 2: the production-size fixture at 2000 + 2000 steps, the scalar-order heatsink iteration, then the schedule-knob probe.
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_full.py tests/test_gpu_dropin.py -m gpu -q --timeout 300 -k "production_size or without_the_avx_macro" > gpurun_out/r02p_tests.log 2>&1
tail -12 gpurun_out/r02p_tests.log | cut -c1-600
bash tools/r02_gpu_o.sh
