#!/bin/bash
# First GPU call of a round: everything written since the last time the repository saw a B200, in one go.
#   gpurun --timeout 1500 -- bash tools/first_gpu_call.sh
# Outputs land in gpurun_out/ (merged back by gpurun).  Nothing here changes the repository.
set -u
mkdir -p gpurun_out
# 1. the legs marked "not yet run on hardware", without the expected-failure marker, each in its own process
python -m pytest tests -m gpu -q --runxfail -p no:cacheprovider \
    -k "thread_per_read or reference_kats or lca_known_answers or build_sections_on_gpu or build_dropin or chunk_rule or shorter_than_the_window" > gpurun_out/new_legs.log 2>&1
echo "new legs rc=$?" | tee -a gpurun_out/new_legs.log
# 1b. memcheck of the thread-per-read kernel on the K2 tests (out-of-bounds shared / global accesses show up here first)
GANON_B200_K2=thread timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
    -k "minimisers_seqan3 or (minimisers_random and 19-31) or (minimisers_random and 4-8)" > gpurun_out/memcheck_k2thread.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/memcheck_k2thread.log
# 2. the whole GPU suite as the driver runs it
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" | tee -a gpurun_out/pytest_gpu.log
# 3. K2 (warp per read) against K2t (thread per read): same bench line, the K2 share is in roofline.other_kernels_ms_per_step
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_c2_k2warp.json 2> gpurun_out/bench_c2_k2warp.err
GANON_B200_K2=thread python bench.py --steps 16 --warmup 3 > gpurun_out/bench_c2_k2thread.json 2> gpurun_out/bench_c2_k2thread.err
GANON_B200_K2=thread python bench.py --workload c4 --steps 16 --warmup 3 > gpurun_out/bench_c4_k2thread.json 2> gpurun_out/bench_c4_k2thread.err
# 4. launch list + one full capture of the thread kernel (never a bench value)
GANON_B200_K2=thread ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k2thread.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
GANON_B200_K2=thread ncu --set full --clock-control none --import-source on -k regex:k_minimisers_thread -c 1 -o gpurun_out/k2thread_full \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/new_legs.log gpurun_out/memcheck_k2thread.log gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for n in ("k2warp", "k2thread"):
    try:
        d = json.loads(open("gpurun_out/bench_c2_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f M reads/s" % (d["value"] / 1e6), "e2e %.1f M" % (d["e2e"]["value"] / 1e6), d["roofline"]["other_kernels_ms_per_step"], "parity", d.get("parity"))
    except Exception as e:
        print(n, "no line:", e)
PY
