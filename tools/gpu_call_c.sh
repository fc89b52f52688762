#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -40 gpurun_out/pytest_gpu.log | cut -c1-600
python bench.py --workload c2 --cli --steps 8 --no-cpu-baseline > gpurun_out/bench_c2_cli.json 2> gpurun_out/bench_c2_cli.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2_cli.json").read().strip().splitlines()[-1])
print("cli", d.get("cli"))
PY
