python -m pytest tests/test_gpu_window.py tests/test_gpu_lthm_step.py -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
