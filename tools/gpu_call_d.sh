#!/bin/bash
# round 2, 8-GPU call: the bench's default N>1 arm (c3 bin-sharded, replicas sub-record) and c5 (256 GiB over 8 GPUs)
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt; free -g | head -2 >> gpurun_out/gpus_n$N.txt
T0=$SECONDS
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_sharded_n$N.json 2> gpurun_out/bench_sharded_n$N.err
echo "bench c3 n=$N rc=$? wall $((SECONDS-T0)) s"; tail -3 gpurun_out/bench_sharded_n$N.err
T0=$SECONDS
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err
echo "bench c5 n=$N rc=$? wall $((SECONDS-T0)) s"; tail -3 gpurun_out/bench_c5_n$N.err
python - <<PY
import json
for f in ("bench_sharded_n$N", "bench_c5_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["parity"])
        print({k: d["config"][k] for k in d["config"] if k.startswith("exchange")})
        print(d["roofline"]["other_kernels_ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
        print(d["e2e"]["last_step_breakdown_ms"])
        for k, v in d.get("extra", {}).items():
            print(k, v.get("value"), v.get("e2e", {}).get("value"), v.get("error"))
    except Exception as e:
        print(f, "no line:", e)
PY
