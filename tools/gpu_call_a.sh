#!/bin/bash
# round 2, 1-GPU call: suite, reference arm + default bench line (c3 + sub-records) with wall times, ncu captures
set -u
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/box.txt; df -h /tmp >> gpurun_out/box.txt; nproc >> gpurun_out/box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/box.txt; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -c avx512 >> gpurun_out/box.txt
T0=$SECONDS; python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_c3.json 2> gpurun_out/ref_c3.err; echo "ref_c3 wall $((SECONDS-T0)) s" | tee gpurun_out/ref_c3.time; echo "ref_c3 wall $((SECONDS-T0)) s" | tee gpurun_out/ref_c3.time
T0=$SECONDS; python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default wall $((SECONDS-T0)) s" | tee gpurun_out/bench_default.time; echo "bench_default wall $((SECONDS-T0)) s" | tee gpurun_out/bench_default.time
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_minimisers_thread<\(int\)2>" -c 1 -o gpurun_out/r02_k2t_c2_full python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k2t.log 2>&1
df -h /tmp | tail -1
python - <<'PY'
import json
for n in ("ref_c3", "bench_default"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, d.get("config", {}).get("workload"), "value %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "parity", d.get("parity"), "extra", {k: (v.get("value"), v.get("e2e", {}).get("value"), v.get("parity"), v.get("error")) for k, v in d.get("extra", {}).items()})
    except Exception as e:
        print(n, "no line:", e)
PY
