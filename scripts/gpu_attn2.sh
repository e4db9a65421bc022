#!/bin/bash
# attention kernels: tests + timing of the tcgen05 variants
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 500 python scripts/gpu_time_kernels.py attn_tc 2>&1 | tail -11
