#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_large.py -q -x --durations=5 2>&1 | tail -25
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py 2>&1 | tail -4
