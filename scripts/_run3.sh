python -m pytest tests/test_gpu_window.py -q -m gpu > gpurun_out/r2_t11.log 2>&1
tail -25 gpurun_out/r2_t11.log
