#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "two_gpus" > gpurun_out/pytest_sharded.log 2>&1
echo "pytest sharded rc=$?"; tail -4 gpurun_out/pytest_sharded.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --extras none --workload c3 > gpurun_out/bench_sharded_n$N.json 2> gpurun_out/bench_sharded_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_sharded_n$N.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["parity"])
print(d["roofline"]["other_kernels_ms_per_step"], d["roofline"]["ms_per_launch"], d["e2e"]["last_step_breakdown_ms"])
PY
