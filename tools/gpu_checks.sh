#!/bin/bash
# What a round's GPU time is spent on, as one script per lease size (outputs under gpurun_out/):
#   gpurun --timeout 1800 -- bash tools/gpu_checks.sh            1 GPU: GPU test suite, reference arm, default bench line
#   gpurun --gpus N --timeout 2400 -- bash tools/gpu_checks.sh N N GPUs: 2-GPU tests (N = 2), bin-sharded bench lines (c3; c5 at N = 8)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
  T0=$SECONDS; python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_c3.json 2> gpurun_out/ref_c3.err; echo "reference arm rc=$? wall $((SECONDS-T0)) s"
  T0=$SECONDS; python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$? wall $((SECONDS-T0)) s"
  ./tools/rowgather_bench > gpurun_out/rowgather.jsonl 2>&1
else
  if [ "$N" = "2" ]; then
    python -m pytest tests -m gpu -q -p no:cacheprovider -k "two_gpus" > gpurun_out/pytest_two_gpus.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 gpurun_out/pytest_two_gpus.log
  fi
  T0=$SECONDS
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_sharded_n$N.json 2> gpurun_out/bench_sharded_n$N.err
  echo "bench c3 bin-sharded n=$N rc=$? wall $((SECONDS-T0)) s"
  if [ "$N" = "8" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --workload c5 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
    echo "bench c5 rc=$?"
  fi
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("config", {}).get("workload"), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "parity", (d.get("parity") or {}).get("identical"))
    except Exception:
        pass
PY
