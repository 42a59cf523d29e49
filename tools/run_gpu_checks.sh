mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_transient.py -x -q 2>&1 | tail -30) > gpurun_out/s4g_transient.log
cat gpurun_out/s4g_transient.log
(timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -8) > gpurun_out/s4g_dropin.log
cat gpurun_out/s4g_dropin.log
