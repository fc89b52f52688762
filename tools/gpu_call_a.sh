#!/bin/bash
# round 2, 1-GPU call: the GPU suite as the driver runs it, then the default bench line with wall time
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -30 gpurun_out/pytest_gpu.log | cut -c1-600
T0=$SECONDS; python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default rc=$? wall $((SECONDS-T0)) s"
tail -3 gpurun_out/bench_default.err
df -h /tmp | tail -1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print(d.get("config", {}).get("workload"), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "parity", d.get("parity"))
    print("tier", d.get("host_resident_tier"))
    for k, v in d.get("extra", {}).items():
        print(k, {kk: v.get(kk) for kk in ("value", "parity", "error", "cli", "gpu_mbp_per_s", "ref_mbp_per_s", "gpu_wall_s", "ref_wall_s", "gpu_stderr_tail")})
except Exception as e:
    print("no line:", e)
PY
