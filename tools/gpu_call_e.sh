#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/rowgather_bench > gpurun_out/rowgather.jsonl 2>&1; cat gpurun_out/rowgather.jsonl
python bench.py --workload c2 --cli --build --steps 8 --no-cpu-baseline > gpurun_out/bench_c2_cli.json 2> gpurun_out/bench_c2_cli.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2_cli.json").read().strip().splitlines()[-1])
c = d.get("cli", {})
print("cli", c.get("reads_per_s"), c.get("classify_s"), c.get("host_pipeline_s")); print("gz", c.get("gz", {}).get("reads_per_s"), c.get("gz", {}).get("host_pipeline_s"))
print("build", d.get("extra", {}).get("ganon_build"))
PY
