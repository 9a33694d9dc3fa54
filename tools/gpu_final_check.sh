#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the default bench arm and its reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 4 --warmup 3 | cut -c1-200
python bench.py | cut -c1-400
