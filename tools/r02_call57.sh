#!/bin/bash
# the streaming regression test (fixed reference), twice
timeout 600 python -m pytest tests/test_gpu_batched.py -m gpu -q -x -k "streaming" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_batched.py -m gpu -q -x -k "bench_shape" 2>&1 | tail -3
