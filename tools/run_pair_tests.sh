#!/bin/bash
# exercise the opt-in 2-CTA multicast GEMM path
GCV_GEMM_PAIR=1 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" "$@"
