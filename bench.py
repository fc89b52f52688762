#!/usr/bin/env python
"""Benchmark of the ganon-classify hot path on B200 (contract: see the task prompt; numbers: BASELINE.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|tiny] [--impl reference]

One "step" = one pass of the hot path (K2 minimisers -> K3 IBF count -> sort of the sparse matches -> K4 rel-filter /
fpr-query / LCA / output lines) over one batch of synthetic 150 bp reads.  `value` = reads/s with the FASTQ batch already in HBM; `e2e` = the same metric through the
C-ABI call gnb_session_classify with HOST (pinned) FASTQ buffers: record indexing, H2D, kernels, D2H and the host
finishing stage (rel-filter, fpr-query, formatting of the `.all` lines) inside the timed region.
N > 1: one process per GPU (torchrun), database replicated, reads sharded -- no data-path collective ("weak").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_sec_150bp"
WORKLOADS = {
    # BASELINE.json configs[1]: 8 GiB flat IBF, 4096 bins, k=19 w=31 h=4, single-end 150 bp
    "c2": dict(bins=4096, bin_size=1 << 24, h=4, k=19, w=31, paired=False, reads_per_step=1 << 21, genome_len=10000, desc="8 GiB flat IBF, 4096 bins, k=19 w=31 h=4, 150 bp single-end"),
    # BASELINE.json configs[2]: 64 GiB flat IBF, 65536 bins, paired
    "c3": dict(bins=65536, bin_size=1 << 23, h=4, k=19, w=31, paired=True, reads_per_step=1 << 18, genome_len=4000, desc="64 GiB flat IBF, 65536 bins, k=19 w=31 h=4, 150 bp paired"),
    # BASELINE.json configs[3]: 3-level HIBF, top 1024 bins (256 merged) -> 256 children of 64 bins (4 merged each) -> 1024
    # grandchildren of 64 bins; 16 + 16 + 8 = 40 GiB; 81 376 user bins, some split over two technical bins
    "c4": dict(hibf=True, top_bins=1024, top_rows=1 << 27, child_bins=64, child_rows=1 << 23, child_merged=4, grand_bins=64, grand_rows=1 << 20, h=4, k=19, w=31,
               paired=False, reads_per_step=1 << 21, genome_len=3000, desc="3-level HIBF, 1024-bin top level, 40 GiB, k=19 w=31 h=4, 150 bp single-end"),
    # the same shape at 1/64 of the rows (self-test of the generator)
    "c4tiny": dict(hibf=True, top_bins=1024, top_rows=1 << 21, child_bins=64, child_rows=1 << 17, child_merged=4, grand_bins=64, grand_rows=1 << 14, h=4, k=19, w=31,
                   paired=False, reads_per_step=1 << 16, genome_len=3000, desc="3-level HIBF, 1024-bin top level, 640 MiB (bench self-test)"),
    # BASELINE.json configs[4]: 256 GiB flat IBF, 65536 bins, bin-sharded over the GPUs (--shard-db; 32 GiB per GPU at N=8)
    "c5": dict(bins=65536, bin_size=1 << 25, h=4, k=19, w=31, paired=True, reads_per_step=1 << 18, genome_len=4000, desc="256 GiB flat IBF, 65536 bins, k=19 w=31 h=4, 150 bp paired"),
    # small stand-in used by the tests of this file
    "tiny": dict(bins=256, bin_size=1 << 16, h=4, k=19, w=31, paired=False, reads_per_step=1 << 14, genome_len=3000, desc="2 MiB flat IBF, 256 bins (bench self-test)"),
}
REL_CUTOFF, REL_FILTER, FPR_QUERY = 0.75, 0.1, 1e-5  # ganon CLI defaults (src/ganon/config.py:603,612,711)
DB_SEED, READ_SEED = 1, 2
CACHE = os.environ.get("GANON_B200_BENCH_DIR", "/tmp/ganon_b200_bench")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ganon-classify")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self.gpu):
                self.rows.append(f)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "power_w_max": max(float(r[3]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------------
# synthetic workload: identical for both arms
# ----------------------------------------------------------------------------------------------------------------------
def target_hashes_for_density(wl, density=0.5):
    """hashes_count such that the declared per-target fpr equals the real one of a filter with this bit density."""
    # (1 - exp(-h*n/m))^h = density^h  ->  n = -ln(1-density) * m / h
    return int(-np.log(1 - density) * wl["bin_size"] / wl["h"])


def hibf_layout(wl):
    """Synthetic 3-level HIBF in the raptor layout: per sub-IBF its technical bins, next_ibf_id and
    ibf_bin_to_filename_position rows; returns also, per user bin, the chain [(ibf, [bins])] from its own IBF up to the
    top level (the user bin's content is inserted along the whole chain, as raptor's merged bins hold their subtree)."""
    bins, rows, nxt, pos = [], [], [], []
    chains = []  # per user bin

    def new_ibf(n_bins, n_rows):
        bins.append(n_bins)
        rows.append(n_rows)
        nxt.append([len(bins) - 1] * n_bins)
        pos.append([0] * n_bins)
        return len(bins) - 1

    def user_bin(ibf, bs, up):
        u = len(chains)
        for b in bs:
            pos[ibf][b] = u
        chains.append([(ibf, bs)] + up)

    top = new_ibf(wl["top_bins"], wl["top_rows"])
    b = 0
    while b < wl["top_bins"]:
        if b % 4 == 0:  # merged bin -> child IBF
            c = new_ibf(wl["child_bins"], wl["child_rows"])
            nxt[top][b], pos[top][b] = c, -1
            up_c = [(top, [b])]
            for cb in range(wl["child_merged"]):
                g = new_ibf(wl["grand_bins"], wl["grand_rows"])
                nxt[c][cb], pos[c][cb] = g, -1
                for gb in range(wl["grand_bins"]):
                    user_bin(g, [gb], [(c, [cb])] + up_c)
            cb = wl["child_merged"]
            user_bin(c, [cb, cb + 1], up_c)  # a user bin split over two technical bins
            for cb in range(wl["child_merged"] + 2, wl["child_bins"]):
                user_bin(c, [cb], up_c)
            b += 1
        elif b % 4 == 1 and (b // 4) % 8 == 0:
            user_bin(top, [b, b + 1], [])  # split
            b += 2
        else:
            user_bin(top, [b], [])
            b += 1
    return bins, rows, nxt, pos, chains


def build_database_hibf(wl, device):
    from ganon_b200 import synth
    from ganon_b200.classify import Database, minimisers_batch

    bins, rows, nxt, pos, chains = hibf_layout(wl)
    names = ["U%d" % u for u in range(len(chains))]
    db = Database.create_hibf(bins, rows, wl["h"], wl["k"], wl["w"], nxt, pos, names, fpr=0.05, device=device)
    db.fill_random(DB_SEED, 1)
    genomes = synth.random_genomes(DB_SEED, len(chains), wl["genome_len"])
    per_ibf = {}  # ibf -> ([hashes], [bins])
    step = 2048
    for g0 in range(0, len(chains), step):
        gs = [genomes[i].tobytes() for i in range(g0, min(g0 + step, len(chains)))]
        hoff, hashes = minimisers_batch(gs, wl["k"], wl["w"], device=device)
        for j in range(len(gs)):
            hs = hashes[int(hoff[j]) : int(hoff[j + 1])]
            for ibf, bs in chains[g0 + j]:
                hl, bl = per_ibf.setdefault(ibf, ([], []))
                hl.append(hs)
                bl.append(np.asarray(bs, dtype=np.uint32)[np.arange(hs.size) % len(bs)])  # split bins share the content
    for ibf, (hl, bl) in per_ibf.items():
        db.emplace(np.concatenate(hl), np.concatenate(bl), ibf_index=ibf)
    return db, genomes


def build_database(wl, device, shard=0, n_shards=1):
    """Flat IBF in HBM: random background bits (density 0.5) OR planted genomes (one per bin).  With n_shards > 1 only
    the bin-word columns of `shard` are created (the same bits as that slice of the whole filter)."""
    from ganon_b200 import synth
    from ganon_b200.classify import Database, minimisers_batch

    if wl.get("hibf"):
        assert n_shards == 1, "bin-block sharding is implemented for flat IBFs"
        return build_database_hibf(wl, device)
    db = Database.create(wl["bins"], wl["bin_size"], wl["h"], wl["k"], wl["w"], device=device, shard=shard, n_shards=n_shards)
    db.fill_random(DB_SEED, 1)
    genomes = synth.random_genomes(DB_SEED, wl["bins"], wl["genome_len"])
    step = 1024
    for g0 in range(0, wl["bins"], step):
        gs = [genomes[i].tobytes() for i in range(g0, min(g0 + step, wl["bins"]))]
        hoff, hashes = minimisers_batch(gs, wl["k"], wl["w"], device=device)
        bins = np.repeat(np.arange(g0, g0 + len(gs), dtype=np.uint32), np.diff(hoff).astype(np.int64))
        db.emplace(hashes, bins)
    n_t = target_hashes_for_density(wl)
    db.set_targets(["T%d" % b for b in range(wl["bins"])], np.arange(wl["bins"], dtype=np.uint32), np.full(wl["bins"], n_t, dtype=np.uint64), n_t)
    return db, genomes


def make_batch(wl, genomes, batch_index, n_reads):
    from ganon_b200 import synth

    m1, m2, _ = synth.reads_from_genomes(READ_SEED + 7919 * batch_index, genomes, n_reads, paired=wl["paired"])
    b1 = synth.fastq_block(m1, first_index=batch_index * n_reads, suffix=b"/1" if wl["paired"] else b"")
    b2 = synth.fastq_block(m2, first_index=batch_index * n_reads, suffix=b"/2") if wl["paired"] else None
    return b1, b2


def bind_to_gpu_numa(dev: int) -> str:
    """N > 1: run this rank on the CPUs of its GPU's NUMA node before any page-locked buffer is allocated, so that the
    host->device copies of the ranks do not all cross the socket interconnect (measured at N=8 without binding: the FASTQ
    copies slow from 12 to 30 ms per step).  Best effort: anything unexpected leaves the process unbound."""
    try:
        import torch

        props = torch.cuda.get_device_properties(dev)
        bus = None
        if hasattr(props, "pci_bus_id") and hasattr(props, "pci_domain_id"):
            bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        else:
            uuid = str(getattr(props, "uuid", ""))
            out = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True, timeout=20).stdout
            for ln in out.splitlines():
                u, b = [x.strip() for x in ln.split(",")]
                if uuid and uuid in u:
                    bus = b.lower()[-12:]
        if not bus:
            return "unbound (no PCI id)"
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "unbound (numa_node unknown)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return "unbound (node %d has no usable CPUs)" % node
        os.sched_setaffinity(0, cpus)
        return "node %d (%d CPUs)" % (node, len(cpus))
    except Exception as e:  # never let placement break the run
        return "unbound (%s)" % str(e)[:60]


def pinned(arr):
    import torch

    t = torch.empty(arr.size, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    t.numpy()[:] = arr
    return t


# ----------------------------------------------------------------------------------------------------------------------
# reference CPU arm: the unmodified ganon-classify (oracle/_ref) on a bounded sample, all host threads
# ----------------------------------------------------------------------------------------------------------------------
def run_reference_binary(ibf_path, fq1, fq2, out_prefix, threads, hibf=False):
    reads = ["-p", fq1 + "," + fq2] if fq2 else ["-r", fq1]
    cmd = [REF_BIN] + (["--hibf"] if hibf else []) + reads + ["-i", ibf_path, "-c", str(REL_CUTOFF), "-d", str(REL_FILTER), "-f", str(FPR_QUERY), "-a", "-o", out_prefix, "-t", str(threads), "--verbose"]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference ganon-classify failed: " + p.stderr[-2000:])
    m = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", p.stderr)
    ml = re.search(r"loading filter\(s\)\s+elapsed \(s\): ([0-9.eE+-]+)", p.stderr)
    return dict(classify_s=float(m.group(1)) if m else wall, load_s=float(ml.group(1)) if ml else 0.0, wall_s=wall)


def reference_threads():
    return max(1, os.cpu_count() or 1)


def ensure_ibf_file(wl_name, db):
    os.makedirs(CACHE, exist_ok=True)
    i = db.info()
    path = os.path.join(CACHE, "%s_seed%d.%s" % (wl_name, DB_SEED, "hibf" if i.is_hibf else "ibf"))
    want = i.device_bytes if i.is_hibf else i.bin_size_bits * i.bin_words * 8
    if not (os.path.exists(path) and os.path.getsize(path) > want):
        free = shutil.disk_usage(CACHE).free
        if free < want * 1.1:
            # make room: the copies written for other workloads are only a cache
            for f in os.listdir(CACHE):
                if f.endswith((".ibf", ".hibf")) and not f.startswith(wl_name + "_seed"):
                    os.remove(os.path.join(CACHE, f))
            free = shutil.disk_usage(CACHE).free
        if free < want * 1.1:
            raise RuntimeError("not enough disk for the reference's copy of the database (%d GiB needed)" % (want >> 30))
        db.save(path)
    return path


def write_sample(wl_name, b1, b2, n_reads, tag):
    os.makedirs(CACHE, exist_ok=True)
    p1 = os.path.join(CACHE, "%s_%s.1.fq" % (wl_name, tag))
    b1.tofile(p1)
    p2 = None
    if b2 is not None:
        p2 = os.path.join(CACHE, "%s_%s.2.fq" % (wl_name, tag))
        b2.tofile(p2)
    return p1, p2


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ganon_b200", choices=["ganon_b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GANON_B200_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--reads-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pool", type=int, default=4, help="distinct read batches cycled through the steps")
    ap.add_argument("--cli", action="store_true", help="also time the drop-in command line (bin/ganon-classify) on a FASTQ file of the pool's batches and the saved database")
    ap.add_argument("--em", action="store_true", help="also time the EM reassignment (SURVEY 8f.1) on the matches of the e2e batches kept in HBM, next to the CPU restatement of src/ganon/reassign.py on the same .all text")
    ap.add_argument("--shard-db", action="store_true", help="bin-shard the database over the GPUs (every rank classifies the same reads on its columns; tuples all-gathered over NCCL): strong scaling")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.reads_per_step:
        wl["reads_per_step"] = args.reads_per_step
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, wl)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"metric": METRIC, "error": "no CUDA device: the hot path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ganon_b200.classify import Session, result_text

    numa = bind_to_gpu_numa(local_rank) if world > 1 else "single process: unbound"
    if args.shard_db:
        return sharded_arm(args, wl, rank, local_rank, world)
    dev = local_rank
    R = wl["reads_per_step"]
    t_setup = time.perf_counter()
    db, genomes = build_database(wl, dev)
    info = db.info()
    pool = max(1, min(args.pool, args.steps + args.warmup))
    # every rank classifies its own reads (read-sharding): batch indices are disjoint across ranks
    blocks = [make_batch(wl, genomes, rank * 1000 + i, R) for i in range(pool)]
    host = [(pinned(b1), pinned(b2) if b2 is not None else None) for b1, b2 in blocks]
    stream = torch.cuda.Stream()
    mk = lambda: Session([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev, cuda_stream=stream.cuda_stream)
    sessions = [mk() for _ in range(pool)]
    for s, (h1, h2) in zip(sessions, host):
        assert s.stage(h1, h2, final=True) == R
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: inputs resident in HBM
    for i in range(max(args.warmup, pool)):
        sessions[i % pool].run_staged()
    sampler = ClockSampler(dev)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    ms_count = ms_min = ms_sort = ms_fin = 0.0
    k3_bytes = launches = minimisers = 0
    for i in range(args.steps):
        r = sessions[(args.warmup + i) % pool].run_staged()
        ms_count += r.ms_count
        ms_min += r.ms_minimiser
        ms_sort += r.ms_sort
        ms_fin += r.ms_finish_device
        k3_bytes += r.count_kernel_bytes
        launches += r.n_kernel_launches
        minimisers += r.n_minimisers
    ev1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    barrier()

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    # the streaming form of the public call: gnb_session_submit / gnb_session_collect (a file reader would do exactly
    # this); every step copies its FASTQ block host->device and reads the step's result back
    e2e_sess = Session([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev)
    _n, cap = e2e_sess.in_flight()

    e2e_prof = {}

    def e2e_loop(n_steps, first):
        h2d = d2h = n_class = n_lines = pending = 0
        last = None
        t_sub = t_col = s_h2d = s_idx = s_job = 0.0
        for i in range(n_steps):
            h1, h2 = host[(first + i) % pool]
            ta = time.perf_counter()
            e2e_sess.submit(h1, h2, final=True)
            t_sub += time.perf_counter() - ta
            pending += 1
            while pending >= cap or (i == n_steps - 1 and pending):
                ta = time.perf_counter()
                r = e2e_sess.collect()
                t_col += time.perf_counter() - ta
                pending -= 1
                h2d += r.h2d_bytes
                d2h += r.d2h_bytes
                n_class += r.n_classified  # the step's result, read on the host
                n_lines += r.all_len[0]
                s_h2d += r.ms_h2d
                s_idx += r.ms_index
                s_job += r.ms_host_index
                last = r
        e2e_prof.update(mean_submit_call_ms=t_sub * 1e3 / n_steps, mean_collect_call_ms=t_col * 1e3 / n_steps, mean_h2d_ms=s_h2d / n_steps, mean_index_ms=s_idx / n_steps, mean_worker_job_ms=s_job / n_steps)
        return h2d, d2h, n_class, n_lines, last

    e2e_loop(max(3, cap + 1), 0)
    barrier()
    t0 = time.perf_counter()
    h2d, d2h, n_class, n_bytes_all, r = e2e_loop(args.steps, 1)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    last = dict(ms_h2d=r.ms_h2d, ms_index=r.ms_index, ms_minimiser=r.ms_minimiser, ms_count=r.ms_count, ms_sort=r.ms_sort, ms_worker_job=r.ms_host_index, ms_finish_device=r.ms_finish_device, levels_on_device=r.levels_on_device, ms_finish_total=r.ms_host_finish, ms_host_merge=r.ms_d2h, ms_submit_to_collect=r.ms_total, batches_in_flight=cap, **e2e_prof)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    clocks = sampler.stop()

    # ------------------------------------------------------------------ EM reassignment from the matches kept in HBM (--em)
    em_line = None
    if args.em and rank == 0:
        # a looser cutoff than the classification default so that many reads have several candidate targets
        em_sess = Session([db], [0.25], [1.0], [1.0], output_all=True, device=dev)
        em_sess.keep_matches(True)
        texts = []
        for h1, h2 in host[:1]:  # one batch: the CPU leg below is pure Python
            r = em_sess.classify(h1, h2, final=True)
            texts.append(result_text(r, "all"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ones, new_rep, em_info = em_sess.reassign(0, 0.0, 10)
        em_ms = (time.perf_counter() - t0) * 1e3
        all_text = b"".join(texts).decode()
        n_lines = all_text.count("\n")
        from oracle import reassign_oracle as RO  # the checker, timed as the CPU leg of this row

        rep_text = em_sess.report(0).decode()
        t0 = time.perf_counter()
        o_ones, o_rep = RO.reassign_texts(rep_text, {"": all_text}, 0.0, 10)
        port_s = time.perf_counter() - t0
        em_line = {"reads_with_matches": ones[""].count(b"\n"), "match_lines": n_lines, "iterations": em_info[""][0], "reads_with_several_matches": em_info[""][1],
                   "gpu_ms": em_ms, "cpu_port_s": port_s, "cpu_port_kind": "oracle/reassign_oracle.py (restatement of src/ganon/reassign.py, 1 thread, from the .all text)",
                   "identical": ones[""].decode() == o_ones[""] and new_rep.decode() == o_rep, "thresholds": "rel-cutoff 0.25 rel-filter 1 fpr-query 1"}
        em_sess.close()

    # ------------------------------------------------------------------ the drop-in command line on files (--cli)
    cli_line = None
    if args.cli and rank == 0 and world == 1:
        ibf_path = ensure_ibf_file(args.workload, db)
        fq1 = os.path.join(CACHE, "%s_cli.1.fq" % args.workload)
        fq2 = os.path.join(CACHE, "%s_cli.2.fq" % args.workload) if wl["paired"] else None
        reps = 4  # the pool's batches four times over: long enough for the fixed costs (buffers, first allocations) to amortise
        with open(fq1, "wb") as f1:
            for _ in range(reps):
                for b1, _b2 in blocks:
                    b1.tofile(f1)
        if fq2:
            with open(fq2, "wb") as f2:
                for _ in range(reps):
                    for _b1, b2 in blocks:
                        b2.tofile(f2)
        n_cli = reps * pool * R * (2 if wl["paired"] else 1)
        reads = ["-p", fq1 + "," + fq2] if fq2 else ["-r", fq1]
        cmd = [sys.executable, os.path.join(ROOT, "bin", "ganon-classify")] + (["--hibf"] if wl.get("hibf") else []) + reads + ["-i", ibf_path, "-c", str(REL_CUTOFF), "-d", str(REL_FILTER), "-f", str(FPR_QUERY), "-a", "-o", os.path.join(CACHE, "cli_out"), "--verbose", "--device", str(dev)]
        t0 = time.perf_counter()
        pr = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        mc = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", pr.stderr)
        ml = re.search(r"loading filter\(s\)\s+elapsed \(s\): ([0-9.eE+-]+)", pr.stderr)
        cs = float(mc.group(1)) if mc else None
        cli_line = {"rc": pr.returncode, "reads": n_cli, "fastq_bytes": os.path.getsize(fq1) + (os.path.getsize(fq2) if fq2 else 0), "classify_s": cs, "load_s": float(ml.group(1)) if ml else None, "wall_s": wall,
                    "reads_per_s": n_cli / cs if cs else None, "all_bytes": os.path.getsize(os.path.join(CACHE, "cli_out.all")) if os.path.exists(os.path.join(CACHE, "cli_out.all")) else None,
                    "stderr_tail": pr.stderr[-300:] if pr.returncode else ""}

    # ------------------------------------------------------------------ CPU baseline + parity on a bounded sample (rank 0, N=1)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, parity = cpu_baseline(args, wl, db, blocks[0], host[0], sessions[0], result_text)
        except Exception as e:  # the bench line must still be printed
            cpu = {"value": None, "unit": "reads/s", "cores": reference_threads(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = (k3_bytes / 1e9) / (ms_count / 1e3) if ms_count > 0 else 0.0
        units = 2 if wl["paired"] else 1  # reads per record
        line = {
            "metric": METRIC,
            "value": world * args.steps * R * units / (dev_ms / 1e3),
            "unit": "reads/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u64",
            "data": "synthetic",
            "config": {
                "workload": args.workload + ": " + wl["desc"],
                "reads_per_step_per_gpu": R * units,
                "db_bytes": int(info.device_bytes),
                "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY),
                "parallelism": "replicated db, reads sharded x%d" % world if world > 1 else "1 gpu",
                "host_placement": numa,
                "l2": "inputs larger than L2: %d distinct %d MB FASTQ batches cycled, %d GiB filter gathered at random" % (pool, blocks[0][0].size * (2 if wl["paired"] else 1) >> 20, int(info.device_bytes) >> 30),
                "timing": "CUDA events on the launch stream around the K steps (max over ranks); wall %.1f ms" % wall_ms,
                "minimisers_per_read": minimisers / max(1, args.steps * R * units),
                "k2_kernel": "warp per read" if os.environ.get("GANON_B200_K2", "").startswith("w") else "thread per read (k2_thread.cuh)",
            },
            "roofline": {
                "kernel": "k_hibf_count" if wl.get("hibf") else "k_ibf_count",
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "peak_source": peak_src,
                "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(args.workload),
                "algorithmic_bytes_per_launch": k3_bytes / max(1, args.steps),
                "ms_per_launch": ms_count / max(1, args.steps),
                "other_kernels_ms_per_step": {"k_minimisers(x2)+scan": ms_min / args.steps, "radix_sort": ms_sort / args.steps, "k_finish(select+scan+write)": ms_fin / args.steps},
            },
            "cpu_baseline": cpu,
            "e2e": {"value": world * args.steps * R * units / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps, "last_step_breakdown_ms": last, "classified_reads_per_step": n_class // args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "setup_s": t_setup,
        }
        if em_line is not None:
            line["em_reassign"] = em_line
        if cli_line is not None:
            line["cli"] = cli_line
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def sharded_arm(args, wl, rank, local_rank, world):
    """--shard-db: the filter is split by bin-word columns over the ranks; every rank stages the same
    batch, runs K2 + K3 on its columns, the sparse tuples are all-gathered in HBM (NCCL) and sorted + finished (K4) on
    every rank.  Total work is fixed as N grows ("strong")."""
    import torch
    import torch.distributed as dist

    from ganon_b200.sharded import ShardedSession

    dev = local_rank
    R = wl["reads_per_step"]
    units = 2 if wl["paired"] else 1
    t_setup = time.perf_counter()
    db, genomes = build_database(wl, dev, shard=rank, n_shards=world)
    info = db.info()
    pool = max(1, min(args.pool, args.steps + args.warmup))
    blocks = [make_batch(wl, genomes, i, R) for i in range(pool)]  # the same reads on every rank
    host = [(pinned(b1), pinned(b2) if b2 is not None else None) for b1, b2 in blocks]
    stream = torch.cuda.Stream()
    mk = lambda: ShardedSession([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev, cuda_stream=stream.cuda_stream)
    sessions = [mk() for _ in range(pool)]
    for s, (h1, h2) in zip(sessions, host):
        assert s.stage(h1, h2, final=True) == R
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, pool)):
            sessions[i % pool].run_levels(prefix_id=1)
        sampler = ClockSampler(dev)
        sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        t0 = time.perf_counter()
        ms_count = ms_min = ms_sort = ms_fin = 0.0
        k3_bytes = launches = minimisers = exchanged = 0
        for i in range(args.steps):
            s = sessions[(args.warmup + i) % pool]
            s.run_levels(prefix_id=1)
            r = s.staged_timings()
            ms_count += r.ms_count
            ms_min += r.ms_minimiser
            ms_sort += r.ms_sort
            ms_fin += r.ms_finish_device
            k3_bytes += r.count_kernel_bytes
            launches += r.n_kernel_launches
            minimisers += r.n_minimisers
            exchanged += s.last_exchanged_bytes
        ev1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([dev_ms, wall_ms, ms_count], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, ms_count_max = float(t[0]), float(t[1]), float(t[2])
        barrier()

        # e2e: host FASTQ blocks through ShardedSession.classify on every rank (H2D of the block, kernels, exchange,
        # K4, result read back), synchronous per step
        e2e = mk()
        for i in range(3):
            e2e.classify(host[i % pool][0], host[i % pool][1], final=True)
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = n_class = 0
        for i in range(args.steps):
            h1, h2 = host[(1 + i) % pool]
            r = e2e.classify(h1, h2, final=True)
            h2d += r.h2d_bytes
            d2h += r.d2h_bytes
            n_class += r.n_classified
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        last = dict(ms_h2d=r.ms_h2d, ms_index=r.ms_index, ms_minimiser=r.ms_minimiser, ms_count=r.ms_count, ms_sort=r.ms_sort, ms_finish_device=r.ms_finish_device, levels_on_device=r.levels_on_device)
        if world > 1:
            t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0])
    clocks = sampler.stop()
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = (k3_bytes / 1e9) / (ms_count / 1e3) if ms_count > 0 else 0.0
        line = {
            "metric": METRIC,
            "value": args.steps * R * units / (dev_ms / 1e3),
            "unit": "reads/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "u64",
            "data": "synthetic",
            "config": {
                "workload": args.workload + ": " + wl["desc"],
                "reads_per_step": R * units,
                "db_bytes_per_gpu": int(info.device_bytes),
                "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY),
                "parallelism": "bin-sharded x%d (bin-word columns %d..%d of %d on rank 0); every rank classifies the same reads, sparse tuples all-gathered in HBM over NCCL, K4 on every rank" % (world, info.shard_word_begin, info.shard_word_end, info.bin_words),
                "l2": "inputs larger than L2: %d distinct %d MB FASTQ batches cycled, filter shard gathered at random" % (pool, blocks[0][0].size * units >> 20),
                "timing": "CUDA events on the launch stream around the K steps (max over ranks); wall %.1f ms" % wall_ms,
                "minimisers_per_read": minimisers / max(1, args.steps * R * units),
                "k2_kernel": "warp per read" if os.environ.get("GANON_B200_K2", "").startswith("w") else "thread per read (k2_thread.cuh)",
                "exchanged_tuple_bytes_per_step": exchanged // max(1, args.steps),
            },
            "roofline": {
                "kernel": "k_ibf_count",
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "peak_source": peak_src,
                "traffic": None,
                "algorithmic_bytes_per_launch": k3_bytes / max(1, args.steps),
                "ms_per_launch": ms_count / max(1, args.steps),
                "ms_per_launch_max_over_ranks": ms_count_max / max(1, args.steps),
                "note": "per GPU (rank 0): its shard's share of every row",
                "other_kernels_ms_per_step": {"k_minimisers(x2)+scan": ms_min / args.steps, "radix_sort": ms_sort / args.steps, "k_finish(select+scan+write)": ms_fin / args.steps},
            },
            "cpu_baseline": None,
            "e2e": {"value": args.steps * R * units / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps, "last_step_breakdown_ms": last, "classified_reads_per_step": n_class // args.steps, "note": "per rank: every rank copies the same block"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": None,
            "setup_s": t_setup,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of one k_ibf_count launch from the committed ncu capture (profiles/)
TRAFFIC_BYTES_PER_LAUNCH = {"c2": 75373784000 + 10292992}  # profiles/r01_k3_ncu_summary.md (2^21 reads per launch)


def cpu_baseline(args, wl, db, block, host_block, sess, result_text):
    """Reference ganon-classify (all host threads) on one batch; also a bit-exact parity check of that batch."""
    if not os.path.exists(REF_BIN):
        raise RuntimeError("oracle/_ref/ganon-classify is not built")
    ibf = ensure_ibf_file(args.workload, db)
    n = min(wl["reads_per_step"], 1 << 20)
    b1, b2 = block
    rec1 = b1.size // wl["reads_per_step"]
    s1 = b1[: n * rec1]
    s2 = b2[: n * (b2.size // wl["reads_per_step"])] if b2 is not None else None
    p1, p2 = write_sample(args.workload, s1, s2, n, "sample")
    out = os.path.join(CACHE, "ref_out")
    threads = reference_threads()
    t = run_reference_binary(ibf, p1, p2, out, threads, hibf=bool(wl.get("hibf")))
    units = 2 if wl["paired"] else 1
    cpu = {"value": n * units / t["classify_s"], "unit": "reads/s", "cores": threads, "kind": "reference", "sample": "%d reads of batch 0; reference's own classifying+printing time %.2f s (filter load %.1f s excluded)" % (n * units, t["classify_s"], t["load_s"])}
    # parity: the same reads through the C ABI
    r = sess.classify(s1, s2, final=True)
    mine = sorted(result_text(r, "all").decode().splitlines())
    with open(out + ".all") as f:
        ref = sorted(l.rstrip("\n") for l in f)
    parity = {"reads": n * units, "all_lines": len(ref), "identical": mine == ref}
    return cpu, parity


def reference_arm(args, wl):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample per step."""
    import torch

    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ganon-classify missing (build it with make -C oracle ref where /root/reference exists)"}))
        return 0
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "the synthetic database is generated on the GPU; no CUDA device here"}))
        return 0
    db, genomes = build_database(wl, 0)  # data generation only; the timed path below is the unmodified binary
    ibf = ensure_ibf_file(args.workload, db)
    db.close()
    n = min(wl["reads_per_step"], 1 << 19)
    threads = reference_threads()
    units = 2 if wl["paired"] else 1
    times = []
    for i in range(args.warmup + args.steps):
        b1, b2 = make_batch(wl, genomes, i % 2, n)
        p1, p2 = write_sample(args.workload, b1, b2, n, "ref%d" % (i % 2))
        t = run_reference_binary(ibf, p1, p2, os.path.join(CACHE, "ref_arm_out"), threads, hibf=bool(wl.get("hibf")))
        if i >= args.warmup:
            times.append(t["classify_s"])
    total = sum(times)
    value = args.steps * n * units / total
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "reads/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": args.workload + ": " + wl["desc"], "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY)},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": "each step = unmodified ganon-classify --threads %d on %d reads (its own classifying+printing time; filter load excluded)" % (threads, n * units)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
