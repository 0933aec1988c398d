#!/bin/bash
set -u
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -k "logits" 2>&1 | tail -n 4
