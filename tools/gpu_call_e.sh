#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "long_reads or build_file_hashes or several_filters or paged" > gpurun_out/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -5 gpurun_out/pytest_new.log | cut -c1-400
python bench.py --workload tiny --build --steps 3 --no-cpu-baseline > gpurun_out/bench_tiny_build.json 2> gpurun_out/bench_tiny_build.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_tiny_build.json").read().strip().splitlines()[-1])
print("build", d.get("extra", {}).get("ganon_build"))
PY
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_hibf_count_narrow" -c 1 -o gpurun_out/r02_k3h_round0_full python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k3h.log 2>&1
tail -2 gpurun_out/ncu_k3h.log | cut -c1-300
