#!/bin/bash
set -u
timeout 300 python scripts/gpu_e2e_probe.py 2>&1 | grep " ms"
