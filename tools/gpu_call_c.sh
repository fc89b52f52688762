#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -40 gpurun_out/pytest_gpu.log | cut -c1-700
