#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
