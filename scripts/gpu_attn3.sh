#!/bin/bash
ATC_MODE=31 timeout 300 python scripts/gpu_time_kernels.py attn_tc 2>&1 | tail -16
