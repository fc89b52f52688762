#!/bin/bash
# round 2, multi-GPU call: the library-level sharded exchange (NCCL) -- parity test, then the bench's default N>1 arm
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "sharded_two_gpus" > gpurun_out/pytest_sharded.log 2>&1
  echo "pytest sharded rc=$?" | tee -a gpurun_out/pytest_sharded.log
  tail -15 gpurun_out/pytest_sharded.log
fi
T0=$SECONDS
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_sharded_n$N.json 2> gpurun_out/bench_sharded_n$N.err
echo "bench n=$N rc=$? wall $((SECONDS-T0)) s"
tail -5 gpurun_out/bench_sharded_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_sharded_n$N.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["parity"])
    print(d["config"]["parallelism"]); print({k: d["config"][k] for k in d["config"] if k.startswith("exchange")})
    print(d["roofline"]["other_kernels_ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
    print(d["e2e"])
    for k, v in d.get("extra", {}).items():
        print(k, v.get("value"), v.get("e2e", {}).get("value"), v.get("error"))
except Exception as e:
    print("no line:", e)
PY
