#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/rowgather_bench > gpurun_out/rowgather.jsonl 2>&1; cat gpurun_out/rowgather.jsonl
