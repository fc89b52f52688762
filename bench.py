#!/usr/bin/env python
"""Benchmark of the ganon-classify hot path on B200 (contract: see the task prompt; numbers: BASELINE.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|tiny] [--impl reference]

Default workload: c3 = BASELINE.json configs[2] (64 GiB flat IBF, 65 536 bins, paired 150 bp reads), the configuration the
north-star targets are quoted on.  N = 1: the c3 line, with c2 (configs[1]) and c4 (configs[3], HIBF) as sub-records under
"extra".  N > 1 (torchrun): the c3 filter bin-sharded over the ranks ("strong" scaling; the N = 1 line is the same
workload on one shard), with the replicated-database / read-sharded arm as a sub-record.

One "step" = one pass of the hot path (K2 minimisers -> K3 IBF count -> sort of the sparse matches -> K4 rel-filter /
fpr-query / LCA / output lines) over one batch of synthetic 150 bp reads.  `value` = reads/s with the FASTQ batch already in HBM; `e2e` = the same metric through the
C-ABI call gnb_session_classify with HOST (pinned) FASTQ buffers: record indexing, H2D, kernels, D2H and the host
finishing stage (rel-filter, fpr-query, formatting of the `.all` lines) inside the timed region.
`--impl reference`: the unmodified reference binary (oracle/_ref/ganon-classify, all host threads), ONE invocation over
steps x n reads of the same workload, on a database file written on the CPU by oracle/synthdb (the same bits the GPU arm
generates in HBM) -- that process never loads libganon_b200.so.
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
    "c2": dict(bins=4096, bin_size=1 << 24, h=4, k=19, w=31, paired=False, reads_per_step=1 << 21, cpu_sample=1 << 20, ref_reads_per_step=1 << 19, genome_len=10000, desc="8 GiB flat IBF, 4096 bins, k=19 w=31 h=4, 150 bp single-end"),
    # BASELINE.json configs[2]: 64 GiB flat IBF, 65536 bins, paired
    "c3": dict(bins=65536, bin_size=1 << 23, h=4, k=19, w=31, paired=True, reads_per_step=1 << 18, cpu_sample=1 << 17, ref_reads_per_step=1 << 15, genome_len=4000, desc="64 GiB flat IBF, 65536 bins, k=19 w=31 h=4, 150 bp paired"),
    # BASELINE.json configs[3]: 3-level HIBF, top 1024 bins (256 merged) -> 256 children of 64 bins (4 merged each) -> 1024
    # grandchildren of 64 bins; 16 + 16 + 8 = 40 GiB; 81 376 user bins, some split over two technical bins
    "c4": dict(hibf=True, top_bins=1024, top_rows=1 << 27, child_bins=64, child_rows=1 << 23, child_merged=4, grand_bins=64, grand_rows=1 << 20, h=4, k=19, w=31,
               paired=False, reads_per_step=1 << 21, cpu_sample=1 << 18, ref_reads_per_step=1 << 16, genome_len=3000, desc="3-level HIBF, 1024-bin top level, 40 GiB, k=19 w=31 h=4, 150 bp single-end"),
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
def _pick_reference_binary():
    """The unmodified reference compiled by oracle/Makefile where /root/reference exists.  Its sources do not travel, so it
    cannot be rebuilt -march=native on the bench box; two builds travel instead and the one for the widest instruction set
    this host runs is used: x86-64-v4 (AVX-512) when /proc/cpuinfo lists it and the binary starts, else AVX2 + BMI2."""
    base = os.path.join(ROOT, "oracle", "_ref", "ganon-classify")
    common = "g++ -std=c++20 -O3 -DNDEBUG %s (oracle/Makefile; the reference's sources do not travel to the bench box, so -march=native is not possible there)"
    v4 = base + ".v4"
    try:
        flags = set()
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    flags = set(ln.split(":", 1)[1].split())
                    break
        if os.path.exists(v4) and {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags and not os.environ.get("GANON_B200_REF_AVX2"):
            if subprocess.run([v4, "--version"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=20).returncode == 0:
                return v4, common % "-march=x86-64-v4"
    except Exception:
        pass
    return base, common % "-mavx2 -mbmi2 -mpopcnt"


REF_BIN, REF_FLAGS = _pick_reference_binary()


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


def make_room(need_bytes, keep=()):
    """The scratch directory holds database copies for the reference binary (up to 64 GiB each) that are only a cache:
    when the disk is short, delete the ones not named in `keep`, largest first (the GPU boxes have ~80 GB free)."""
    os.makedirs(CACHE, exist_ok=True)
    want = int(need_bytes * 1.05) + (2 << 30)
    if shutil.disk_usage(CACHE).free >= want:
        return
    cand = [os.path.join(CACHE, f) for f in os.listdir(CACHE) if f.endswith((".ibf", ".hibf", ".part")) and os.path.join(CACHE, f) not in keep]
    for p in sorted(cand, key=os.path.getsize, reverse=True):
        os.remove(p)
        if shutil.disk_usage(CACHE).free >= want:
            return
    if shutil.disk_usage(CACHE).free < need_bytes:
        raise RuntimeError("not enough disk in %s (%d GiB needed)" % (CACHE, need_bytes >> 30))


def ensure_ibf_file(wl_name, db):
    os.makedirs(CACHE, exist_ok=True)
    i = db.info()
    path = os.path.join(CACHE, "%s_seed%d.%s" % (wl_name, DB_SEED, "hibf" if i.is_hibf else "ibf"))
    want = i.device_bytes if i.is_hibf else i.bin_size_bits * i.bin_words * 8
    if not (os.path.exists(path) and os.path.getsize(path) > want):
        make_room(want)
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


def db_file_matches_hbm(db, path, windows=32, words=1 << 14, seed=7):
    """Spot check that the cached database file holds the bits that are in HBM: `windows` random runs of 64-bit words of
    the flat filter's payload (the file may have been written on the CPU by oracle/synthdb in the reference arm)."""
    i = db.info()
    if i.is_hibf:
        return None
    n_words = i.bin_size_bits * i.bin_words
    off0 = os.path.getsize(path) - n_words * 8  # the payload is the tail of the file (formats.py)
    rng = np.random.default_rng(seed)
    ok = True
    with open(path, "rb") as f:
        for _ in range(windows):
            w0 = int(rng.integers(0, max(1, n_words - words)))
            n = int(min(words, n_words - w0))
            f.seek(off0 + w0 * 8)
            want = np.frombuffer(f.read(n * 8), dtype="<u8")
            ok = ok and bool((db.read_words(w0, n) == want).all())
    return {"windows": windows, "words_per_window": words, "identical": ok}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel from a committed `ncu --set full`
# capture: (bytes, reads per launch of that capture, file under profiles/).  Only attached to a line whose launch has the
# same shape; otherwise "traffic" is null.
TRAFFIC_CAPTURES = {
    "c2": (75373784000 + 10292992, 1 << 21, "profiles/r01_k3_ncu_summary.md"),
    "c3": (301541113000 + 4524544, 1 << 19, "profiles/r02_k3_c3_ncu_summary.md"),
    "c3": (301541113000 + 4524544, 1 << 19, "profiles/r02_k3_c3_ncu_summary.md"),
}


# useful GB/s of a pure random row gather by row width (bytes), measured with tools/rowgather_bench.cu on B200
ROW_GATHER_GBPS = {8: 334.0, 32: 1168.5, 64: 2334.8, 128: 4652.3, 256: 6389.4, 512: 6799.0}


def traffic_for(wl_name, reads_per_launch):
    cap = TRAFFIC_CAPTURES.get(wl_name)
    if cap and cap[1] == reads_per_launch:
        return cap[0], cap[2]
    return None, None


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ganon_b200", choices=["ganon_b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GANON_B200_WORKLOAD", ""), choices=[""] + sorted(WORKLOADS), help="default: c3 with c2 / c4 as sub-records (N = 1), c3 bin-sharded (N > 1)")
    ap.add_argument("--extras", default=None, help="comma list of workloads measured as sub-records of the line (default: c2,c4 at N = 1 when --workload is not given; 'none' to skip)")
    ap.add_argument("--reads-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pool", type=int, default=4, help="distinct read batches cycled through the steps")
    ap.add_argument("--cli", action="store_true", help="also time the drop-in command line (bin/ganon-classify) on a FASTQ file of the pool's batches and the saved database (always on for the c2 sub-record of the default run)")
    ap.add_argument("--em", action="store_true", help="also time the EM reassignment (SURVEY 8f.1) on the matches of the e2e batches kept in HBM, next to the CPU restatement of src/ganon/reassign.py on the same .all text")
    ap.add_argument("--paged", action="store_true", help="also measure the host-resident tier: the workload's filter with half of it allowed in HBM (always on for the default c3 line)")
    ap.add_argument("--build", action="store_true", help="also time the drop-in ganon-build on the GPU next to the reference builder (always on for the default N = 1 line)")
    ap.add_argument("--shard-db", action="store_true", help="bin-shard the database over the GPUs (the default for N > 1)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: replicated database, reads sharded (weak scaling) as the headline instead of the bin-sharded arm")
    args = ap.parse_args()
    explicit = bool(args.workload)
    wl_name = args.workload or "c3"
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, wl_name)

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
    ctx = dict(rank=rank, local_rank=local_rank, world=world, numa=bind_to_gpu_numa(local_rank) if world > 1 else "single process: unbound")

    if args.extras is None:
        extras = ["c2", "c4"] if (world == 1 and not explicit) else []
    else:
        extras = [e for e in args.extras.split(",") if e and e != "none"]

    if (world > 1 and not args.replicas) or args.shard_db:
        line = sharded_arm(args, wl_name, ctx)
        if not explicit and world > 1:
            # the other way to use N GPUs (the database fits one GPU): replicas, reads sharded, no data-path collective
            sub = measure(args, "c2", ctx, cpu=False, cli=False, em=False)
            if rank == 0:
                line.setdefault("extra", {})["replicas_c2"] = sub
    else:
        line = measure(args, wl_name, ctx, cpu=not args.no_cpu_baseline, cli=args.cli, em=args.em, paged=args.paged or (not explicit and world == 1))
        for e in extras:
            try:
                sub = measure(args, e, ctx, cpu=not args.no_cpu_baseline, cli=(e == "c2"), em=False)
            except Exception as ex:  # a sub-record must not take the headline with it
                sub = {"error": str(ex)[:300]}
            if rank == 0:
                line.setdefault("extra", {})[e] = sub
    if rank == 0 and world == 1 and (args.build or not explicit):
        try:
            line.setdefault("extra", {})["ganon_build"] = build_leg(local_rank)
        except Exception as ex:
            line.setdefault("extra", {})["ganon_build"] = {"error": str(ex)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def measure(args, wl_name, ctx, cpu, cli, em, paged=False):
    """One workload on every rank's own GPU: database replicated, reads sharded over the ranks (no data-path collective).
    Returns the bench line (rank 0; None elsewhere)."""
    import torch
    import torch.distributed as dist

    from ganon_b200.classify import Session, result_text

    rank, world, dev = ctx["rank"], ctx["world"], ctx["local_rank"]
    wl = dict(WORKLOADS[wl_name])
    if args.reads_per_step:
        wl["reads_per_step"] = args.reads_per_step
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
    hibf_rounds = sessions[(args.warmup + args.steps - 1) % pool].hibf_rounds() if wl.get("hibf") else None
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
    e2e_sess.close()

    # ------------------------------------------------------------------ EM reassignment from the matches kept in HBM (--em)
    em_line = None
    if em and rank == 0:
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

    # ------------------------------------------------------------------ CPU baseline + parity on a bounded sample (rank 0, N=1)
    cpu_line = parity = file_check = None
    if rank == 0 and world == 1 and cpu:
        try:
            cpu_line, parity, file_check = cpu_baseline(wl_name, wl, db, blocks[0], sessions[0], result_text)
        except Exception as e:  # the bench line must still be printed
            cpu_line = {"value": None, "unit": "reads/s", "cores": reference_threads(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}

    # ------------------------------------------------------------------ the drop-in command line on files (--cli)
    cli_line = None
    if cli and rank == 0 and world == 1:
        try:
            cli_line = cli_leg(wl_name, wl, db, blocks, pool, R, dev)
        except Exception as e:
            cli_line = {"error": str(e)[:300]}

    # ------------------------------------------------------------------ host-resident tier: the same filter with half
    # of it allowed in HBM, the rest streamed from page-locked host memory per batch
    paged_line = None
    if paged and rank == 0 and world == 1 and not wl.get("hibf"):
        try:
            for s in sessions:
                s.close()
            sessions = []
            paged_line = paged_leg(wl, db, host, dev, Session, result_text)
        except Exception as e:
            paged_line = {"error": str(e)[:300]}

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = (k3_bytes / 1e9) / (ms_count / 1e3) if ms_count > 0 else 0.0
        units = 2 if wl["paired"] else 1  # reads per record
        traffic, traffic_src = traffic_for(wl_name, R * units)
        line = {
            "metric": METRIC,
            "value": world * args.steps * R * units / (dev_ms / 1e3),
            "unit": "reads/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True,
            # N = 1 is the one-shard end of the bin-sharded strong-scaling series; replicas scale weakly
            "scaling": "weak" if world > 1 else "strong",
            "vs_baseline": None,
            "dtype": "u64",
            "data": "synthetic",
            "config": {
                "workload": wl_name + ": " + wl["desc"],
                "reads_per_step_per_gpu": R * units,
                "db_bytes": int(info.device_bytes),
                "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY),
                "parallelism": "replicated db, reads sharded x%d" % world if world > 1 else "1 gpu",
                "host_placement": ctx["numa"],
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
                "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": k3_bytes / max(1, args.steps),
                "ms_per_launch": ms_count / max(1, args.steps),
                "other_kernels_ms_per_step": {"k_minimisers+scan": ms_min / args.steps, "radix_sort": ms_sort / args.steps, "k_finish(select+scan+write)": ms_fin / args.steps},
            },
            "cpu_baseline": cpu_line,
            "e2e": {"value": world * args.steps * R * units / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps, "last_step_breakdown_ms": last, "classified_reads_per_step": n_class // args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "db_file_check": file_check,
            "setup_s": t_setup,
        }
        if hibf_rounds:
            # an HIBF traversal gathers rows of very different widths; the copy peak is the wrong yardstick for narrow rows:
            # tools/rowgather_bench.cu measures what random rows of each width can reach on this GPU
            rowb = [wl["top_bins"] // 8, wl["child_bins"] // 8, wl["grand_bins"] // 8]
            line["roofline"]["rounds"] = []
            for i, (ms_r, by_r, it_r) in enumerate(hibf_rounds):
                rb = rowb[min(i, len(rowb) - 1)]
                gp = ROW_GATHER_GBPS.get(rb)
                gb = by_r / 1e9 / (ms_r / 1e3) if ms_r > 0 else 0.0
                line["roofline"]["rounds"].append({"round": i, "items": it_r, "row_bytes": rb, "ms": ms_r, "algorithmic_bytes": by_r, "achieved_GBps": gb,
                                                   "row_gather_roofline_GBps": gp, "frac_of_row_gather_roofline": gb / gp if gp else None})
            line["roofline"]["row_gather_roofline_source"] = "profiles/r02_rowgather_bench.jsonl (tools/rowgather_bench.cu on B200: random rows from 32 GiB, 16 loads in flight per lane)"
        if em_line is not None:
            line["em_reassign"] = em_line
        if cli_line is not None:
            line["cli"] = cli_line
        if paged_line is not None:
            line["host_resident_tier"] = paged_line
    for s in sessions:
        s.close()
    db.close()
    del host, blocks
    torch.cuda.empty_cache()
    return line


def sharded_arm(args, wl_name, ctx):
    """N > 1 default: the filter is split by bin-word columns over the ranks.  Every rank sees the same
    batch, runs K2 + K3 on its columns; the sparse tuples are exchanged in HBM inside the library (NCCL) and sorted +
    finished (K4) on every rank.  Total work is fixed as N grows ("strong").  e2e: submit / collect with sliced ingest
    (each rank copies 1/N of the FASTQ block over its own PCIe link, slices all-gathered over NVLink)."""
    import torch
    import torch.distributed as dist

    from ganon_b200.classify import Session, result_text
    from ganon_b200.sharded import ShardedSession, make_comm

    rank, world, dev = ctx["rank"], ctx["world"], ctx["local_rank"]
    wl = dict(WORKLOADS[wl_name])
    if args.reads_per_step:
        wl["reads_per_step"] = args.reads_per_step
    R = wl["reads_per_step"]
    units = 2 if wl["paired"] else 1
    t_setup = time.perf_counter()
    db, genomes = build_database(wl, dev, shard=rank, n_shards=world)
    info = db.info()
    comm = make_comm(dev)
    pool = max(1, min(args.pool, args.steps + args.warmup))
    blocks = [make_batch(wl, genomes, i, R) for i in range(pool)]  # the same reads on every rank
    host = [(pinned(b1), pinned(b2) if b2 is not None else None) for b1, b2 in blocks]
    stream = torch.cuda.Stream()
    mk = lambda **kw: ShardedSession([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev, comm=comm, **kw)
    sessions = [mk(cuda_stream=stream.cuda_stream) for _ in range(pool)]
    for s, (h1, h2) in zip(sessions, host):
        assert s.stage(h1, h2, final=True) == R
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_(t, op):
        if world > 1:
            dist.all_reduce(t, op=op)

    # ------------------------------------------------------------------ value: inputs resident in HBM
    for i in range(max(args.warmup, pool)):
        sessions[i % pool].run_staged()
    sampler = ClockSampler(dev)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    ms_count = ms_min = ms_sort = ms_fin = ms_xch = 0.0
    k3_bytes = launches = minimisers = exchanged = 0
    for i in range(args.steps):
        r = sessions[(args.warmup + i) % pool].run_staged()
        ms_count += r.ms_count
        ms_min += r.ms_minimiser
        ms_sort += r.ms_sort
        ms_fin += r.ms_finish_device
        ms_xch += r.ms_exchange
        k3_bytes += r.count_kernel_bytes
        launches += r.n_kernel_launches
        minimisers += r.n_minimisers
        exchanged += r.exchanged_bytes
    ev1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms, wall_ms, ms_count, ms_xch], device="cuda", dtype=torch.float64)
    reduce_(t, dist.ReduceOp.MAX)
    dev_ms, wall_ms, ms_count_max, ms_xch_max = (float(x) for x in t)
    barrier()

    # ------------------------------------------------------------------ e2e: host FASTQ blocks through submit / collect
    e2e = mk(sliced_ingest=True)
    _n, cap = e2e.in_flight()
    e2e_prof = {}

    def e2e_loop(n_steps, first):
        h2d = d2h = n_class = pending = 0
        last = None
        s_xch = 0.0
        for i in range(n_steps):
            h1, h2 = host[(first + i) % pool]
            e2e.submit(h1, h2, final=True)
            pending += 1
            while pending >= cap or (i == n_steps - 1 and pending):
                r = e2e.collect()
                pending -= 1
                h2d += r.h2d_bytes
                d2h += r.d2h_bytes
                n_class += r.n_classified
                s_xch += r.ms_exchange
                last = r
        e2e_prof.update(mean_exchange_ms=s_xch / n_steps)
        return h2d, d2h, n_class, last

    e2e_loop(max(3, cap + 1), 0)
    barrier()
    t0 = time.perf_counter()
    h2d, d2h, n_class, r = e2e_loop(args.steps, 1)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    last = dict(ms_h2d=r.ms_h2d, ms_index=r.ms_index, ms_minimiser=r.ms_minimiser, ms_count=r.ms_count, ms_exchange=r.ms_exchange, ms_sort=r.ms_sort, ms_finish_device=r.ms_finish_device,
                levels_on_device=r.levels_on_device, ms_submit_to_collect=r.ms_total, batches_in_flight=cap, **e2e_prof)
    t = torch.tensor([e2e_ms, float(h2d)], device="cuda", dtype=torch.float64)
    tm = t.clone()
    reduce_(tm, dist.ReduceOp.MAX)
    reduce_(t, dist.ReduceOp.SUM)
    e2e_ms, h2d_all = float(tm[0]), int(t[1])
    clocks = sampler.stop()

    # ------------------------------------------------------------------ parity: the sharded result of batch 0 against the
    # unsharded session on one GPU (which the N = 1 line checks against the reference binary), or -- when the whole
    # filter does not fit one GPU -- against the oracle on a sample (rows of the filter regenerated on the CPU)
    res = e2e.classify(host[0][0], host[0][1], final=True)  # collective: every rank
    mine = result_text(res, "all")
    parity = None
    if rank == 0:
        full_bytes = wl["bin_size"] * ((wl["bins"] + 63) // 64) * 8
        free, _total = torch.cuda.mem_get_info(dev)
        try:
            if full_bytes + (12 << 30) < free and not os.environ.get("GANON_B200_BENCH_ORACLE_PARITY"):
                parity = parity_against_unsharded(wl, dev, host[0], mine, R * units)
            else:
                parity = parity_against_oracle_sample(wl, dev, genomes, blocks[0], mine)
        except Exception as ex:
            parity = {"identical": None, "error": str(ex)[:300]}
    e2e.close()

    line = None
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
                "workload": wl_name + ": " + wl["desc"],
                "reads_per_step": R * units,
                "db_bytes_per_gpu": int(info.device_bytes),
                "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY),
                "parallelism": "bin-sharded x%d (bin-word columns %d..%d of %d on rank 0); every rank classifies the same reads on its columns, sparse tuples exchanged in HBM inside libganon_b200 (NCCL %d), K4 on every rank" % (world, info.shard_word_begin, info.shard_word_end, info.bin_words, comm.nccl_version()),
                "host_placement": ctx["numa"],
                "l2": "inputs larger than L2: %d distinct %d MB FASTQ batches cycled, filter shard gathered at random" % (pool, blocks[0][0].size * units >> 20),
                "timing": "CUDA events on the launch stream around the K steps (max over ranks); wall %.1f ms" % wall_ms,
                "minimisers_per_read": minimisers / max(1, args.steps * R * units),
                "k2_kernel": "warp per read" if os.environ.get("GANON_B200_K2", "").startswith("w") else "thread per read (k2_thread.cuh)",
                "exchanged_tuple_bytes_per_step": exchanged // max(1, args.steps),
                "exchange_ms_per_step": ms_xch / args.steps,
                "exchange_ms_per_step_max_over_ranks": ms_xch_max / args.steps,
                "exchange_share_of_step": (ms_xch_max / args.steps) / (dev_ms / args.steps),
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
                "traffic_source": None,
                "algorithmic_bytes_per_launch": k3_bytes / max(1, args.steps),
                "ms_per_launch": ms_count / max(1, args.steps),
                "ms_per_launch_max_over_ranks": ms_count_max / max(1, args.steps),
                "note": "per GPU (rank 0): its shard's share of every row",
                "other_kernels_ms_per_step": {"k_minimisers+scan": ms_min / args.steps, "tuple_exchange": ms_xch / args.steps, "radix_sort": ms_sort / args.steps, "k_finish(select+scan+write)": ms_fin / args.steps},
            },
            "cpu_baseline": None,
            "e2e": {"value": args.steps * R * units / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d_all // args.steps, "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "last_step_breakdown_ms": last, "classified_reads_per_step": n_class // args.steps, "note": "sliced ingest: h2d bytes summed over the ranks (each copies 1/N of the block); d2h of rank 0 (every rank reads the full result back)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "setup_s": t_setup,
        }
    for s in sessions:
        s.close()
    db.close()
    comm.close()
    del host, blocks
    torch.cuda.empty_cache()
    return line


def parity_against_unsharded(wl, dev, host_block, sharded_all_text, n_reads):
    from ganon_b200.classify import Session, result_text

    db, _g = build_database(wl, dev)
    s = Session([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev)
    r = s.classify(host_block[0], host_block[1], final=True)
    want = result_text(r, "all")
    out = {"reads": n_reads, "all_lines": want.count(b"\n"), "identical": sorted(want.splitlines()) == sorted(sharded_all_text.splitlines()),
           "byte_identical": want == sharded_all_text, "against": "the unsharded session (whole filter on rank 0's GPU) on the same batch; that path is checked against the reference binary in the N = 1 line"}
    s.close()
    db.close()
    return out


def parity_against_oracle_sample(wl, dev, genomes, block, sharded_all_text, n_sample=192):
    """The whole filter fits neither one GPU nor (for the reference binary) the host: check the first n_sample records of
    the batch against the oracle (oracle/ganon_oracle.c, the pinned CPU restatement).  The oracle reads the filter through
    a sparse anonymous mapping of its full size in which only the rows those reads touch are materialised, regenerated on
    the CPU with the numpy statements of the device generators (ganon_b200/synth.py: background words + planted bits)."""
    import ctypes as C
    import mmap

    from ganon_b200 import synth
    from ganon_b200.classify import minimisers_batch
    from oracle import oracle as O

    k, w, h, bins, bin_size = wl["k"], wl["w"], wl["h"], wl["bins"], wl["bin_size"]
    bw = (bins + 63) // 64
    b1, b2 = block
    rec1 = b1.size // wl["reads_per_step"]
    rec2 = b2.size // wl["reads_per_step"] if b2 is not None else 0

    def records(b, rec):
        out = []
        for i in range(n_sample):
            lines = bytes(b[i * rec : (i + 1) * rec]).split(b"\n")
            out.append((lines[0][1:], lines[1]))
        return out

    r1 = records(b1, rec1)
    r2 = records(b2, rec2) if b2 is not None else None
    reads = [(r1[i][0], r1[i][1], r2[i][1] if r2 else None) for i in range(n_sample)]
    hashes = [O.read_hashes(s1, s2, k, w) for _id, s1, s2 in reads]
    allh = np.unique(np.concatenate([x for x in hashes if x is not None]))
    need = np.unique(synth.ibf_rows(allh, h, bin_size).reshape(-1))
    n_words = bin_size * bw
    libc = C.CDLL(None, use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    MAP_NORESERVE = 0x4000
    addr = libc.mmap(None, n_words * 8, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | MAP_NORESERVE, -1, 0)
    if addr in (None, C.c_void_p(-1).value):
        raise RuntimeError("sparse mapping of the filter failed (errno %d)" % C.get_errno())
    data = np.ctypeslib.as_array((C.c_uint64 * n_words).from_address(addr))
    try:
        for r in need.tolist():
            data[r * bw : (r + 1) * bw] = synth.random_words(DB_SEED, 1, bin_size, bw, bins, row0=r, rows=1)
        # planted bits on those rows: every genome's minimisers (K2 on the GPU -- data generation, as in build_database)
        step = 4096
        for g0 in range(0, bins, step):
            gs = [genomes[i].tobytes() for i in range(g0, min(g0 + step, bins))]
            hoff, hs = minimisers_batch(gs, k, w, device=dev)
            gbin = np.repeat(np.arange(g0, g0 + len(gs), dtype=np.uint64), np.diff(hoff).astype(np.int64))
            rows = synth.ibf_rows(hs, h, bin_size)
            for i in range(h):
                sel = np.isin(rows[i], need)
                if sel.any():
                    idx = (rows[i][sel] * np.uint64(bw) + (gbin[sel] >> np.uint64(6))).astype(np.int64)
                    np.bitwise_or.at(data, idx, np.uint64(1) << (gbin[sel] & np.uint64(63)))
        oibf = O.OracleIBF(bins, bin_size, h, data)
        n_t = target_hashes_for_density(wl)
        fpr = O.lib().go_target_fpr(bin_size, h, n_t, n_t)
        filt = O.OracleFilter(oibf, ["T%d" % b for b in range(bins)], [[b] for b in range(bins)], [fpr] * bins, REL_CUTOFF, k, w)
        want = O.all_lines(O.classify_level([filt], reads, REL_FILTER, FPR_QUERY))
    finally:
        del data
        libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
        libc.munmap(addr, n_words * 8)
    ids = {r[0] for r in reads}
    mine = sorted(ln for ln in sharded_all_text.decode().splitlines() if ln.split("\t", 1)[0].encode() in ids)
    return {"reads": n_sample * (2 if b2 is not None else 1), "all_lines": len(want), "identical": mine == want, "rows_materialised": int(need.size),
            "against": "oracle (CPU restatement, pinned to the reference) on the first %d records of batch 0; filter rows regenerated on the CPU" % n_sample}


def paged_leg(wl, db, host, dev, Session, result_text):
    """The workload's filter with an HBM budget of half its size: column pages, part resident, part streamed per batch."""
    import torch

    units = 2 if wl["paired"] else 1
    mk = lambda: Session([db], [REL_CUTOFF], [REL_FILTER], [FPR_QUERY], output_all=True, device=dev)
    s = mk()
    want = result_text(s.classify(host[0][0], host[0][1], final=True), "all")
    s.close()
    full = int(db.info().device_bytes)
    budget = full // 2
    t0 = time.perf_counter()
    db.page_out(budget)
    t_out = time.perf_counter() - t0
    info = db.info()
    s = mk()
    got = result_text(s.classify(host[0][0], host[0][1], final=True), "all")
    # large batches amortise the stream (every batch streams the non-resident pages once): all pool blocks, four times
    # over, as one block per step
    reps = 4
    big1 = pinned(np.tile(np.concatenate([h1.numpy() for h1, _h2 in host]), reps))
    big2 = pinned(np.tile(np.concatenate([h2.numpy() for _h1, h2 in host]), reps)) if host[0][1] is not None else None
    n_big = reps * len(host) * wl["reads_per_step"] * units
    r = s.classify(big1, big2, final=True)
    torch.cuda.synchronize()
    steps = 2
    t0 = time.perf_counter()
    streamed = 0
    for _ in range(steps):
        r = s.classify(big1, big2, final=True)
        streamed += r.h2d_bytes
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s.close()
    return {"filter_bytes": full, "hbm_budget_bytes": budget, "pages": int(info.n_pages), "resident_pages": int(info.n_resident_pages), "hbm_bytes": int(info.device_bytes), "host_bytes": int(info.host_bytes),
            "resident_fraction": 1.0 - info.host_bytes / full, "page_out_s": t_out, "reads_per_step": n_big, "steps": steps, "ms_per_step": dt / steps * 1e3, "value": steps * n_big / dt, "unit": "reads/s",
            "h2d_bytes_per_step": streamed // steps, "h2d_GBps": streamed / dt / 1e9, "ms_count_last": r.ms_count,
            "bound": "PCIe: every batch streams the non-resident pages once; K3 of the resident pages and of the page before overlaps the copies",
            "parity": {"reads": wl["reads_per_step"] * units, "identical": got == want, "against": "the same session on the filter whole in HBM (byte comparison of the .all text)"}}


def build_leg(dev, n_targets=128, genome_len=4_000_000):
    """`ganon-build`: the drop-in builder on the GPU next to the unmodified reference builder with all host
    threads, same FASTA files and input table; parity on what is invariant in the reference (IBF parameters, per-target hash
    counts, number of bins per target) and on classification: both filters classify the same reads identically."""
    from ganon_b200 import formats, synth
    from ganon_b200.classify import Database, Session, result_text

    ref_build = os.path.join(ROOT, "oracle", "_ref", "ganon-build")
    d = os.path.join(CACHE, "build")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    make_room(n_targets * genome_len * 3)
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    table = []
    for t in range(n_targets):
        g = acgt[rng.integers(0, 4, size=genome_len, dtype=np.uint8)]
        lines = np.empty((genome_len // 80, 81), dtype=np.uint8)
        lines[:, :80] = g[: genome_len // 80 * 80].reshape(-1, 80)
        lines[:, 80] = ord("\n")
        path = os.path.join(d, "g%04d.fna" % t)
        with open(path, "wb") as f:
            f.write(b">contig%d\n" % t)
            f.write(lines.tobytes())
        table.append("%s\tT%d" % (path, t))
    tab = os.path.join(d, "input.tsv")
    with open(tab, "w") as f:
        f.write("\n".join(table) + "\n")
    n_bp = n_targets * (genome_len // 80 * 80)
    common = ["--input-file", tab, "--kmer-size", "19", "--window-size", "31", "--max-fp", "0.05", "--hash-functions", "4", "--verbose"]
    out_gpu, out_ref = os.path.join(d, "gpu.ibf"), os.path.join(d, "ref.ibf")
    os.makedirs(os.path.join(d, "tmp_gpu"), exist_ok=True)  # both builders want an existing folder
    os.sync()  # the genomes were just written: let the write-back finish (measured: 1.4 - 4.3 s for the same build without it)
    t0 = time.perf_counter()
    pg = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "ganon-build")] + common + ["--output-file", out_gpu, "--tmp-output-folder", os.path.join(d, "tmp_gpu") + "/", "--device", str(dev)],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    t_gpu = time.perf_counter() - t0
    own = lambda text: (lambda m: float(m.group(1)) if m else None)(re.search(r"processed .* in ([0-9.eE+-]+) seconds", text))
    line = {"targets": n_targets, "bases": n_bp, "gpu_rc": pg.returncode, "gpu_wall_s": t_gpu, "gpu_own_s": own(pg.stderr), "gpu_mbp_per_s": n_bp / 1e6 / t_gpu,
            "gpu_stderr_tail": pg.stderr[-200:] if pg.returncode else "", "note": "wall = whole process (interpreter + CUDA start included); own = the builder's reported time"}
    if os.path.exists(ref_build) and pg.returncode == 0:
        os.makedirs(os.path.join(d, "tmp_ref"), exist_ok=True)
        os.sync()
        t0 = time.perf_counter()
        pr = subprocess.run([ref_build] + common + ["--output-file", out_ref, "--tmp-output-folder", os.path.join(d, "tmp_ref") + "/", "--threads", str(reference_threads())],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        t_ref = time.perf_counter() - t0
        line.update(ref_rc=pr.returncode, ref_wall_s=t_ref, ref_own_s=own(pr.stderr), ref_mbp_per_s=n_bp / 1e6 / t_ref, ref_threads=reference_threads(), ref_stderr_tail=pr.stderr[-200:] if pr.returncode else "")
        if pr.returncode == 0:
            a, b = formats.read_ibf(out_gpu, load_data=False), formats.read_ibf(out_ref, load_data=False)
            same_cfg = (a.ibf.bins, a.ibf.bin_size, a.ibf.hash_funs, a.kmer_size, a.window_size, a.max_hashes_bin) == (b.ibf.bins, b.ibf.bin_size, b.ibf.hash_funs, b.kmer_size, b.window_size, b.max_hashes_bin)
            same_counts = dict(a.hashes_count) == dict(b.hashes_count)
            nb = lambda x: sorted((t, sum(1 for _b, tt in x.bin_map if tt == t)) for t in dict(x.hashes_count))
            # classification with both filters: reads sampled from the genomes
            reads = []
            for i in range(20000):
                t = int(rng.integers(0, n_targets))
                with open(os.path.join(d, "g%04d.fna" % t), "rb") as f:
                    f.seek(len(b">contig%d\n" % t) + int(rng.integers(0, genome_len // 80 - 4)) * 81)
                    s = f.read(81 * 3).replace(b"\n", b"")[:150]
                reads.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
            fq = b"".join(reads)
            outs = []
            for path in (out_gpu, out_ref):
                dbx = Database.open(path, device=dev)
                sx = Session([dbx], [0.5], [0.1], [1.0], output_all=True, device=dev)
                outs.append(sorted(result_text(sx.classify(fq, final=True), "all").decode().splitlines()))
                sx.close()
                dbx.close()
            line["parity"] = {"ibf_parameters_equal": same_cfg, "per_target_hash_counts_equal": same_counts, "bins_per_target_equal": nb(a) == nb(b),
                              "classification_equal": outs[0] == outs[1], "classified_lines": len(outs[0]),
                              "note": "bit layout inside a target's bins follows hash-table iteration order in the reference: compared on the invariants"}
    shutil.rmtree(d, ignore_errors=True)
    return line


def cli_leg(wl_name, wl, db, blocks, pool, R, dev):
    """bin/ganon-classify file to file: the pool's batches (four times over) as a FASTQ file, the saved database."""
    ibf_path = ensure_ibf_file(wl_name, db)
    fq1 = os.path.join(CACHE, "%s_cli.1.fq" % wl_name)
    fq2 = os.path.join(CACHE, "%s_cli.2.fq" % wl_name) if wl["paired"] else None
    reps = 4  # long enough for the fixed costs (buffers, first allocations) to amortise
    make_room(reps * sum(b1.size + (b2.size if b2 is not None else 0) for b1, b2 in blocks) * 1.5, keep=(ibf_path,))
    with open(fq1, "wb") as f1:
        for _ in range(reps):
            for b1, _b2 in blocks:
                b1.tofile(f1)
    if fq2:
        with open(fq2, "wb") as f2:
            for _ in range(reps):
                for _b1, b2 in blocks:
                    b2.tofile(f2)
    os.sync()  # the files were just written: let the write-back finish, or it competes with the run for the host's memory system
    n_cli = reps * pool * R * (2 if wl["paired"] else 1)
    reads = ["-p", fq1 + "," + fq2] if fq2 else ["-r", fq1]
    cmd = [sys.executable, os.path.join(ROOT, "bin", "ganon-classify")] + (["--hibf"] if wl.get("hibf") else []) + reads + ["-i", ibf_path, "-c", str(REL_CUTOFF), "-d", str(REL_FILTER), "-f", str(FPR_QUERY), "-a", "-o", os.path.join(CACHE, "cli_out"), "--verbose", "--device", str(dev)]
    # two consecutive runs, the second one reported: the classification phase lasts well under a second, and the first run
    # starts on a GPU that sat idle while this benchmark wrote its input files (clocks ramp up during it; measured: its
    # staging calls take 2-3 x longer than in any following run, whatever the host settings)
    first_cs = None
    for attempt in range(2):
        t0 = time.perf_counter()
        pr = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        mc = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", pr.stderr)
        if attempt == 0:
            first_cs = float(mc.group(1)) if mc else None
            if pr.returncode != 0:
                break
    ml = re.search(r"loading filter\(s\)\s+elapsed \(s\): ([0-9.eE+-]+)", pr.stderr)
    cs = float(mc.group(1)) if mc else None
    out = {"first_run_classify_s": first_cs, "rc": pr.returncode, "reads": n_cli, "fastq_bytes": os.path.getsize(fq1) + (os.path.getsize(fq2) if fq2 else 0), "classify_s": cs, "load_s": float(ml.group(1)) if ml else None, "wall_s": wall,
           "reads_per_s": n_cli / cs if cs else None, "all_bytes": os.path.getsize(os.path.join(CACHE, "cli_out.all")) if os.path.exists(os.path.join(CACHE, "cli_out.all")) else None,
           "stderr_tail": pr.stderr[-300:] if pr.returncode else ""}
    mh = re.search(r"host pipeline \(s\): ([^\n]*)", pr.stderr)
    out["host_pipeline_s"] = mh.group(1) if mh else None
    if os.environ.get("GANON_B200_BENCH_CLI_VARIANTS"):
        # diagnosis aid: the same command under other host settings
        out["variants"] = {}
        for tag, env in (("io_threads_4", {"GANON_B200_IO_THREADS": "4"}), ("io_threads_8", {"GANON_B200_IO_THREADS": "8"}), ("sync_spin", {"GANON_B200_SYNC": "spin"}),
                         ("block_256MiB", {"GANON_B200_BLOCK_BYTES": str(256 << 20)})):
            pv = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, **env))
            mv = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", pv.stderr)
            mp = re.search(r"host pipeline \(s\): ([^\n]*)", pv.stderr)
            out["variants"][tag] = {"classify_s": float(mv.group(1)) if mv else None, "host_pipeline_s": mp.group(1) if mp else None}
    # the same reads as ordinary single-member gzip files (what sequencers / archives deliver): the library inflates them
    # with all host threads (csrc/gzstream.cpp)
    try:
        gz1 = fq1 + ".gz"
        gz2 = fq2 + ".gz" if fq2 else None
        t0 = time.perf_counter()
        write_single_member_gzip(gz1, [b1 for b1, _b2 in blocks])
        if gz2:
            write_single_member_gzip(gz2, [b2 for _b1, b2 in blocks])
        t_gz = time.perf_counter() - t0
        os.sync()
        n_gz = pool * R * (2 if wl["paired"] else 1)
        reads_gz = ["-p", gz1 + "," + gz2] if gz2 else ["-r", gz1]
        cmd_gz = [c for c in cmd]
        i = cmd_gz.index("-p" if fq2 else "-r")
        cmd_gz[i : i + 2] = reads_gz
        t0 = time.perf_counter()
        pr = subprocess.run(cmd_gz, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        mc = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", pr.stderr)
        cs = float(mc.group(1)) if mc else None
        out["gz"] = {"rc": pr.returncode, "reads": n_gz, "gz_bytes": os.path.getsize(gz1) + (os.path.getsize(gz2) if gz2 else 0), "fastq_bytes": sum(b1.size + (b2.size if b2 is not None else 0) for b1, b2 in blocks),
                     "classify_s": cs, "wall_s": wall, "reads_per_s": n_gz / cs if cs else None, "compress_s": t_gz, "host_threads": reference_threads(),
                     "kind": "single-member gzip (one deflate stream per file, level 6)", "stderr_tail": pr.stderr[-300:] if pr.returncode else ""}
        mh = re.search(r"host pipeline \(s\): ([^\n]*)", pr.stderr)
        out["gz"]["host_pipeline_s"] = mh.group(1) if mh else None
        if os.environ.get("GANON_B200_BENCH_CLI_VARIANTS"):
            out["gz"]["variants"] = {}
            for tag, env in (("io_threads_8", {"GANON_B200_IO_THREADS": "8"}), ("io_threads_10", {"GANON_B200_IO_THREADS": "10"}), ("io_threads_12", {"GANON_B200_IO_THREADS": "12"}),
                             ("io_threads_14", {"GANON_B200_IO_THREADS": "14"}), ("io_threads_16", {"GANON_B200_IO_THREADS": "16"}), ("default_again", {}),
                             ("block_32MiB", {"GANON_B200_BLOCK_BYTES": str(32 << 20)}), ("block_16MiB_t14", {"GANON_B200_BLOCK_BYTES": str(16 << 20), "GANON_B200_IO_THREADS": "14"})):
                pv = subprocess.run(cmd_gz, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, **env))
                mv = re.search(r"classifying\+printing elapsed \(s\): ([0-9.eE+-]+)", pv.stderr)
                mp = re.search(r"host pipeline \(s\): ([^\n]*)", pv.stderr)
                out["gz"]["variants"][tag] = {"classify_s": float(mv.group(1)) if mv else None, "host_pipeline_s": mp.group(1) if mp else None}
        for p in (gz1, gz2):
            if p and os.path.exists(p):
                os.remove(p)
    except Exception as e:
        out["gz"] = {"error": str(e)[:300]}
    for p in (fq1, fq2, os.path.join(CACHE, "cli_out.all")):
        if p and os.path.exists(p):
            os.remove(p)
    return out


def write_single_member_gzip(path, arrays, level=6, piece=8 << 20):
    """One gzip member whose deflate stream is compressed in pieces on all host threads (each piece ends with a full flush,
    the way pigz -i writes): an ordinary .gz for every reader, made quickly."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    views = []
    for a in arrays:
        mv = memoryview(np.ascontiguousarray(a)).cast("B")
        views += [mv[o : o + piece] for o in range(0, len(mv), piece)]

    def comp(iv):
        i, v = iv
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return co.compress(v) + co.flush(zlib.Z_FINISH if i == len(views) - 1 else zlib.Z_FULL_FLUSH), zlib.crc32(v), len(v)

    crc = 0
    total = 0
    with ThreadPoolExecutor(reference_threads()) as ex, open(path, "wb") as f:
        f.write(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03")
        for data, c, n in ex.map(comp, enumerate(views)):
            f.write(data)
            crc = _crc32_combine(crc, c, n)
            total += n
        f.write(struct.pack("<II", crc & 0xFFFFFFFF, total & 0xFFFFFFFF))


def _crc32_combine(crc1, crc2, len2):
    """zlib's crc32_combine (not exposed by Python): CRC of a concatenation from the CRCs of its parts."""

    def times(mat, vec):
        s = 0
        i = 0
        while vec:
            if vec & 1:
                s ^= mat[i]
            vec >>= 1
            i += 1
        return s

    def square(mat):
        return [times(mat, mat[n]) for n in range(32)]

    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]
    even = square(odd)
    odd = square(even)
    while True:
        even = square(odd)
        if len2 & 1:
            crc1 = times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = square(even)
        if len2 & 1:
            crc1 = times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def cpu_baseline(wl_name, wl, db, block, sess, result_text):
    """Reference ganon-classify (all host threads) on a sample of batch 0; also a bit-exact parity check of that sample
    and a spot check of the database file against HBM."""
    if not os.path.exists(REF_BIN):
        raise RuntimeError("oracle/_ref/ganon-classify is not built")
    ibf = ensure_ibf_file(wl_name, db)
    file_check = db_file_matches_hbm(db, ibf)
    n = min(wl["reads_per_step"], wl.get("cpu_sample", 1 << 20))
    b1, b2 = block
    rec1 = b1.size // wl["reads_per_step"]
    s1 = b1[: n * rec1]
    s2 = b2[: n * (b2.size // wl["reads_per_step"])] if b2 is not None else None
    p1, p2 = write_sample(wl_name, s1, s2, n, "sample")
    out = os.path.join(CACHE, "ref_out")
    threads = reference_threads()
    t = run_reference_binary(ibf, p1, p2, out, threads, hibf=bool(wl.get("hibf")))
    units = 2 if wl["paired"] else 1
    cpu = {"value": n * units / t["classify_s"], "unit": "reads/s", "cores": threads, "kind": "reference", "flags": REF_FLAGS,
           "sample": "%d reads of batch 0; reference's own classifying+printing time %.2f s (filter load %.1f s excluded)" % (n * units, t["classify_s"], t["load_s"])}
    # parity: the same reads through the C ABI
    r = sess.classify(s1, s2, final=True)
    mine = sorted(result_text(r, "all").decode().splitlines())
    with open(out + ".all") as f:
        ref = sorted(l.rstrip("\n") for l in f)
    parity = {"reads": n * units, "all_lines": len(ref), "identical": mine == ref, "against": "unmodified reference binary on the same reads and database file"}
    for p in (p1, p2, out + ".all"):
        if p and os.path.exists(p):
            os.remove(p)
    return cpu, parity, file_check


# ----------------------------------------------------------------------------------------------------------------------
# --impl reference: this process never imports ganon_b200._lib / loads libganon_b200.so
# ----------------------------------------------------------------------------------------------------------------------
def reference_db_file(wl_name, wl):
    """The workload's database as a file, written on the CPU by oracle/synthdb (flat IBFs).  A copy left in the cache by
    an earlier run (either arm) is reused: both writers produce the same bits (db_file_check in the GPU arm's line)."""
    from ganon_b200 import synth  # numpy only

    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "%s_seed%d.%s" % (wl_name, DB_SEED, "hibf" if wl.get("hibf") else "ibf"))
    genomes = None
    if wl.get("hibf"):
        if not os.path.exists(path):
            raise RuntimeError("no CPU writer for the synthetic HIBF: run the GPU arm with --workload %s once (it saves the file)" % wl_name)
        return path, None, "cached copy saved by the GPU arm"
    genomes = synth.random_genomes(DB_SEED, wl["bins"], wl["genome_len"])
    want = wl["bin_size"] * ((wl["bins"] + 63) // 64) * 8
    if os.path.exists(path) and os.path.getsize(path) > want:
        return path, genomes, "cached copy"
    make_room(want + genomes.size)
    tool = os.path.join(ROOT, "oracle", "synthdb")
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "synthdb"])
    gfile = os.path.join(CACHE, "%s_genomes.bin" % wl_name)
    genomes.tofile(gfile)
    t0 = time.perf_counter()
    out = subprocess.run([tool, path, gfile, str(wl["bins"]), str(wl["bin_size"]), str(wl["h"]), str(wl["k"]), str(wl["w"]), str(wl["genome_len"]), str(DB_SEED), str(target_hashes_for_density(wl)), str(reference_threads())],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    os.remove(gfile)
    if out.returncode != 0:
        raise RuntimeError("oracle/synthdb failed: " + out.stderr[-300:])
    return path, genomes, "written by oracle/synthdb in %.0f s (words xor sum planted: %s)" % (time.perf_counter() - t0, out.stdout.strip())


def reference_arm(args, wl_name):
    """--impl reference: the reference's own CPU implementation (unmodified binary), all host threads: one invocation over
    steps x n reads of the workload, filter load paid once, its own classifying+printing time / steps."""
    wl = dict(WORKLOADS[wl_name])
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ganon-classify missing (build it with make -C oracle ref where /root/reference exists)"}))
        return 0
    try:
        ibf, genomes, how = reference_db_file(wl_name, wl)
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:300]}))
        return 0
    if genomes is None:
        from ganon_b200 import synth

        genomes = synth.random_genomes(DB_SEED, len(hibf_layout(wl)[4]), wl["genome_len"])
    n = min(wl["reads_per_step"], wl.get("ref_reads_per_step", 1 << 19))  # records per step: a bounded sample of the workload
    threads = reference_threads()
    units = 2 if wl["paired"] else 1
    p1 = os.path.join(CACHE, "%s_refarm.1.fq" % wl_name)
    p2 = os.path.join(CACHE, "%s_refarm.2.fq" % wl_name) if wl["paired"] else None
    f1 = open(p1, "wb")
    f2 = open(p2, "wb") if p2 else None
    for i in range(args.steps):
        b1, b2 = make_batch(wl, genomes, i, n)  # same generator and seeds as the GPU arm's batches
        b1.tofile(f1)
        if f2:
            b2.tofile(f2)
    f1.close()
    if f2:
        f2.close()
    t = run_reference_binary(ibf, p1, p2, os.path.join(CACHE, "ref_arm_out"), threads, hibf=bool(wl.get("hibf")))
    for p in (p1, p2, os.path.join(CACHE, "ref_arm_out.all")):
        if p and os.path.exists(p):
            os.remove(p)
    total = t["classify_s"]
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
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": wl_name + ": " + wl["desc"], "thresholds": "rel-cutoff %.2f rel-filter %.2f fpr-query %g" % (REL_CUTOFF, REL_FILTER, FPR_QUERY),
                   "reads_per_step": n * units, "database_file": how, "invocations": 1,
                   "timing": "the binary's own classifying+printing time over steps x n reads / steps; filter load %.1f s paid once and excluded; no separate warm-up pass (one process)" % t["load_s"]},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "reference", "flags": REF_FLAGS,
                         "sample": "one run of the unmodified ganon-classify --threads %d over %d steps x %d reads (wall %.1f s incl. filter load)" % (threads, args.steps, n * units, t["wall_s"])},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    with open("/proc/self/maps") as f:  # evidence for the record: no product library in this process
        line["native_so_loaded"] = sorted({ln.split("/")[-1].strip() for ln in f if "ganon_b200" in ln and ".so" in ln})
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
