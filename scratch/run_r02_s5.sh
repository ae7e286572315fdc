#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmark_parity.py tests/test_gpu_dd_local.py -m gpu -q -x 2>&1 | tail -12 > $O/r02g_tests.log; tail -6 $O/r02g_tests.log
