#!/bin/bash
timeout 500 python -m pytest tests/test_gpu_large.py -q -x 2>&1 | tail -5
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py 2>&1 | tail -5
timeout 600 python tools/gpu_time_predictions.py 25 600 > gpurun_out/r2_predictions_study.json 2> gpurun_out/r2_predictions_study.err
tail -c 600 gpurun_out/r2_predictions_study.json; tail -5 gpurun_out/r2_predictions_study.err
