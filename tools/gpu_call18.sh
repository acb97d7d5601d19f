#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_fullsize_gpu.py -x -q ) > gpurun_out/c18_pytest_full.log 2>&1
tail -25 gpurun_out/c18_pytest_full.log | cut -c1-400
