#!/bin/bash
# first GPU pass of the large-capacity paths (row f3): the new tests, then the whole GPU suite
timeout 500 python -m pytest tests/test_gpu_large.py -q -x --durations=8 2>&1 | tail -40
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py 2>&1 | tail -5
