"""Host arithmetic of `ganon-build`: the choice of the IBF parameters from the per-target minimiser
counts.  This is the part of the build that is not a kernel -- K2 (minimisers) and `k_emplace` (insertion) already exist
in the library; a GPU `ganon-build` is this module in front of them.

Mirrors `optimal_hashes` and its helpers (src/ganon-build/GanonBuild.cpp:290-618) and `true_false_positive` (382-412).
All quantities are IEEE doubles evaluated with libm in the reference's order, so the chosen configuration is the
reference's (pinned against the reference binary by tests/test_build_config_cpu.py).
"""
from __future__ import annotations

import gzip
import math
import os
import re
import sys
import time
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

MIB_BITS = 8388608  # bits per "megabyte" of --filter-size


@dataclass
class IBFParams:
    """IBFConfig (src/utils/include/utils/IBFConfig.hpp:18-40) as chosen by the build."""

    n_bins: int = 0
    max_hashes_bin: int = 0
    hash_functions: int = 0
    bin_size_bits: int = 0
    max_fp: float = 0.0
    true_max_fp: float = 0.0
    true_avg_fp: float = 0.0

    @property
    def technical_bins(self) -> int:
        return padded_bins(self.n_bins)

    @property
    def filter_bits(self) -> int:
        return self.technical_bins * self.bin_size_bits


# The reference computes in IEEE doubles and never traps: log(0) = -inf, x / 0 = +-inf or nan, and a non-finite value cast to
# uint64_t comes out as 2^63 with the x86-64 code gcc emits.  Degenerate requests (--max-fp 1, filters of a few bits) rely on
# that to end in "No valid sequences to build" instead of an exception, so the helpers below follow the same rules.
def _u64(x: float) -> int:
    """C++ conversion of a double to uint64_t: truncation; non-finite or out-of-range values give 2^63 (cvttsd2si)."""
    if x != x or x in (math.inf, -math.inf) or x >= 18446744073709551616.0:
        return 1 << 63
    if x <= -1.0:
        return int(x) % (1 << 64)
    return int(x)


def _log(x: float) -> float:
    return math.log(x) if x > 0 else (-math.inf if x == 0 else math.nan)


def _exp(x: float) -> float:
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def _div(a: float, b: float) -> float:
    if b != 0:
        return a / b
    if a == 0 or a != a:
        return math.nan
    return math.copysign(math.inf, a) * math.copysign(1.0, b)


def _pow(a: float, b: float) -> float:
    try:
        return math.pow(a, b)
    except OverflowError:
        return math.inf
    except ValueError:
        return math.nan


def _ceil(x: float) -> float:
    return float(math.ceil(x)) if math.isfinite(x) else x


def bin_bits_for_fp(max_fp: float, n_hashes: int, hash_functions: int = 0) -> int:
    """Bits of one bin for `n_hashes` elements at false-positive rate `max_fp` (GanonBuild.cpp:290-305): with the optimal
    number of hash functions when none is given, else for that number."""
    if hash_functions == 0:
        return _u64(_ceil(_div(n_hashes * _log(max_fp), _log(1.0 / math.pow(2, math.log(2))))))
    return _u64(_ceil(n_hashes * _div(-hash_functions, _log(1 - _exp(_div(_log(max_fp), hash_functions))))))


def pick_hash_functions(bin_size_bits: int, n_hashes: int, requested: int, limit: int = 5) -> int:
    """Number of hash functions (GanonBuild.cpp:308-333): the requested one, or ln2 * bits/elements truncated to uint8,
    clamped to 1..limit (0 and values above the limit become the limit)."""
    h = requested
    if h == 0:
        h = _u64(math.log(2) * _div(bin_size_bits, float(n_hashes))) & 0xFF  # static_cast<uint8_t>
    if h > limit or h == 0:
        h = limit
    return h


def split_bins(counts: Iterable[int], n_hashes: int) -> int:
    """Bins needed when no bin holds more than n_hashes elements (GanonBuild.cpp:336-347)."""
    return sum(_u64(_ceil(_div(c, float(n_hashes)))) for c in counts)


def padded_bins(n_bins: int) -> int:
    """Technical bins: the next multiple of 64 (GanonBuild.cpp:365-371)."""
    return _u64(math.ceil(n_bins / 64.0) * 64)


def bloom_fp(bin_size_bits: int, hash_functions: int, n_hashes: int) -> float:
    """Theoretical false-positive rate of one bin (GanonBuild.cpp:373-380)."""
    return _pow(1 - _exp(_div(-hash_functions, _div(bin_size_bits, float(n_hashes)))), hash_functions)


def split_correction(max_split_bins: int, max_fp: float, hash_functions: int, n_hashes: int) -> float:
    """Growth of a bin that keeps the rate of a target spread over `max_split_bins` bins at max_fp
    (multiple testing; GanonBuild.cpp:350-362)."""
    target_fpr = 1.0 - _exp(_div(_log(1.0 - max_fp), max_split_bins))
    grown = bin_bits_for_fp(target_fpr, n_hashes, hash_functions)
    base = bin_bits_for_fp(max_fp, n_hashes, hash_functions)
    return _div(float(grown), base)


def real_fp(counts: Iterable[int], max_hashes_bin: int, bin_size_bits: int, hash_functions: int) -> Tuple[float, float]:
    """(highest, average) false-positive rate over the targets with their split bins (GanonBuild.cpp:382-412).  The
    average is summed in the order given (the reference sums in hash-map order: equal up to the last bits)."""
    highest = total = 0.0
    n = 0
    for c in counts:
        n_bins_target = _u64(math.ceil(c / float(max_hashes_bin)))
        if n_bins_target == 0:
            # a target without hashes (every sequence shorter than --min-length): the reference evaluates pow(x, 0) = 1,
            # i.e. a rate of 0 that still counts in the average
            fp = 0.0
        else:
            n_hashes_bin = _u64(math.ceil(c / float(n_bins_target)))
            fp = 1.0 - _pow(1.0 - bloom_fp(bin_size_bits, hash_functions, n_hashes_bin), n_bins_target)
        highest = max(highest, fp)
        total += fp
        n += 1
    return highest, total / float(n)


def choose_ibf_params(hashes_count: Dict[str, int], max_fp: float = 0.05, filter_size: float = 0.0, hash_functions: int = 0,
                      mode: str = "avg", max_hash_functions: int = 5) -> IBFParams:
    """`optimal_hashes` (GanonBuild.cpp:428-618) followed by `true_false_positive`: simulate every bin capacity from the
    largest target downwards in steps of 100 elements, and keep the capacity with the best harmonic mean of
    (filter size or false positive) and (number of bins), each relative to its minimum over the simulations."""
    counts = list(hashes_count.values())
    largest = max(counts) if counts else 0
    if largest == 0:
        return IBFParams()
    sims: List[Tuple[int, int, int, float]] = []  # (n_hashes, n_bins, filter_size_bits, fp)
    min_filter = min_bins = 0
    min_fp = 1.0
    step = min(100, largest)
    n = largest + 1
    while n > step:
        cap = n - 1
        bins = split_bins(counts, cap)
        if filter_size:
            bits = _u64(_div(filter_size, float(padded_bins(bins))) * MIB_BITS)
            h = pick_hash_functions(bits, cap, hash_functions, max_hash_functions)
        elif hash_functions == 0:
            bits = bin_bits_for_fp(max_fp, cap)
            h = pick_hash_functions(bits, cap, 0, max_hash_functions)
        else:
            h = pick_hash_functions(0, cap, hash_functions, max_hash_functions)
            bits = bin_bits_for_fp(max_fp, cap, h)
        most_splits = _u64(math.ceil(largest / float(cap)))
        fp, filter_bits = 0.0, 0
        if filter_size:
            fp = 1 - _pow(1.0 - bloom_fp(bits, h, cap), most_splits)
            min_fp = min(min_fp, fp)
        else:
            per_split = _u64(math.ceil(largest / float(most_splits)))
            approx = min(bloom_fp(bits, h, per_split), max_fp)
            rate = split_correction(most_splits, approx, h, cap)
            bits = _u64(bits * rate)
            filter_bits = (bits * padded_bins(bins)) & 0xFFFFFFFFFFFFFFFF
            if filter_bits == 0 or math.isinf(rate):  # GanonBuild.cpp:541-542
                break
            if filter_bits < min_filter or min_filter == 0:
                min_filter = filter_bits
        sims.append((cap, bins, filter_bits, fp))
        if bins < min_bins or min_bins == 0:
            min_bins = bins
        n -= step
        if step == 0:
            break
    # weights of the two ratios in the mean (the special modes tilt or drop one of them)
    tilt = 0.5 if mode in ("smaller", "faster") else 0.0 if mode in ("smallest", "fastest") else 1.0
    w_var = tilt if mode in ("smaller", "smallest") else 1.0
    w_bins = tilt if mode in ("faster", "fastest") else 1.0
    best = IBFParams()
    best_score = 0.0
    for cap, bins, filter_bits, fp in sims:
        # (0 / 0 = nan when every simulated rate underflows to 0: the first simulation then wins, as in the reference)
        var_ratio = _div(fp, min_fp) if filter_size else _div(filter_bits, float(min_filter))
        bins_ratio = _div(bins, float(min_bins))
        score = (1 + math.pow(tilt, 2)) * _div(var_ratio * bins_ratio, (w_var * var_ratio) + (w_bins * bins_ratio))
        if score < best_score or best_score == 0:
            best_score = score
            if filter_size:
                best.bin_size_bits = _u64((filter_size / float(padded_bins(bins))) * MIB_BITS)
                best.max_fp = fp
            else:
                best.bin_size_bits = filter_bits // padded_bins(bins)
                best.max_fp = max_fp
            best.max_hashes_bin = cap
            best.n_bins = bins
            best.hash_functions = pick_hash_functions(best.bin_size_bits, cap, hash_functions, max_hash_functions)
    if best.n_bins:
        best.true_max_fp, best.true_avg_fp = real_fp(counts, best.max_hashes_bin, best.bin_size_bits, best.hash_functions)
    return best


def bin_layout(hashes_count: Dict[str, int], params: IBFParams) -> List[Tuple[str, int, int]]:
    """`create_bin_map_hash` (GanonBuild.cpp:619-653) for targets in the order given: per technical bin
    (target, first hash index, last hash index) -- a target's hashes are spread evenly over its bins."""
    out: List[Tuple[str, int, int]] = []
    for target, count in hashes_count.items():
        n_bins_target = _u64(math.ceil(count / float(params.max_hashes_bin)))
        if n_bins_target == 0:
            continue  # no hashes, no bins (the loop below would not run in the reference either)
        per_bin = min(_u64(math.ceil(count / float(n_bins_target))), params.max_hashes_bin)
        for i in range(n_bins_target):
            first = i * per_bin
            if first >= count:
                break
            out.append((target, first, min(first + per_bin, count) - 1))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# `ganon-build` (GanonBuild::run, GanonBuild.cpp:752-923): input table -> minimisers per target -> parameters -> filter
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class GanonBuildConfig:
    """Field for field GanonBuild::Config (src/ganon-build/include/ganon-build/Config.hpp:14-31)."""

    input_file: str = ""
    output_file: str = ""
    tmp_output_folder: str = ""
    max_fp: float = 0.05
    filter_size: float = 0.0
    kmer_size: int = 19
    window_size: int = 31
    hash_functions: int = 0
    mode: str = "avg"
    min_length: int = 0
    threads: int = 1
    verbose: bool = False
    quiet: bool = False
    device: int = 0  # not in the reference

    def _err(self, msg: str) -> bool:
        if not self.quiet:
            print(msg, file=sys.stderr)
        return False

    def validate(self) -> bool:
        """Config::validate (Config.hpp:33-107): same checks, same messages."""
        if not self.input_file:
            return self._err("--input-file is mandatory")
        if not os.path.exists(self.input_file):
            return self._err("--input-file not found: " + self.input_file)
        if os.path.getsize(self.input_file) == 0:
            return self._err("--input-file is empty: " + self.input_file)
        if not self.output_file:
            return self._err("--output-file is mandatory")
        if self.tmp_output_folder and not os.path.exists(self.tmp_output_folder):
            return self._err("--tmp-output-folder not found")
        if self.hash_functions > 5:
            return self._err("--hash-functions must be <=5")
        if self.filter_size == 0 and self.max_fp == 0:
            return self._err("--max-fp or --filter-size is mandatory")
        if self.filter_size > 0:
            self.max_fp = 0
        if self.window_size < self.kmer_size:
            return self._err("--window-size has to be >= --kmer-size")
        if self.mode not in ("avg", "smaller", "smallest", "faster", "fastest"):
            return self._err("Invalid --mode")
        if self.kmer_size > 32:
            return self._err("--kmer-size has to be <= 32")
        return True


def parse_input_table(path: str, quiet: bool = False) -> Tuple[Dict[str, List[str]], int]:
    """`parse_input_file` (GanonBuild.cpp:86-137): tab-separated `file [<tab> target]`; the target defaults to the file
    name.  Returns ({target: [files]}, number of missing / empty files).  Targets keep the order of first appearance
    (the reference keeps them in a hash map; the order only decides which bins a target gets)."""
    targets: Dict[str, List[str]] = {}
    invalid = 0
    with open(path) as fh:
        for line in fh.read().split("\n"):
            if not line:
                continue
            fields = line.split("\t")
            if len(fields) > 1 and fields[-1] == "":
                fields.pop()  # std::getline on the fields: nothing follows a trailing tab
            f = fields[0]
            if not os.path.exists(f) or os.path.getsize(f) == 0:
                if not quiet:
                    print("WARNING: input file not found/empty: " + f, file=sys.stderr)
                invalid += 1
                continue
            if len(fields) == 1:
                targets.setdefault(os.path.basename(f), []).append(f)
            elif len(fields) == 2:
                targets.setdefault(fields[1], []).append(f)
    return targets, invalid


_WS = b" \t\n\v\f\r"
_WS_DIGITS = _WS + b"0123456789"
_LEGAL = b"ABCDGHKMNRSTVWYUabcdghkmnrstvwyu"  # dna15 + U (nucleotide_base.hpp:147-168)
_NEXT_ID = re.compile(rb"[>;]")


def read_sequences(path: str) -> Optional[List[bytes]]:
    """Sequences of a FASTA / FASTQ file, plain or gzip, as seqan3::sequence_file_input<dna4_traits> yields them to
    count_hashes (GanonBuild.cpp:205-226); None = parse error, the reference then drops the whole file (its hashes are
    counted after the loop over the records, GanonBuild.cpp:239-247).  The record rules are those of the library's host
    reader (csrc/reads.cpp: format_fasta.hpp:150-330, format_fastq.hpp:105-267), which tests/test_reader_cpu.py pins against
    the reference binary; tests/test_build_cpu.py pins this restatement against the reference builder."""
    with open(path, "rb") as f:
        magic = f.read(2)
    data = gzip.open(path, "rb").read() if magic == b"\x1f\x8b" else open(path, "rb").read()
    out: List[bytes] = []
    n, p = len(data), 0
    if data[:1] in (b">", b";"):
        while p < n:
            if data[p : p + 1] not in (b">", b";"):
                return None  # "Expected to be on beginning of ID"
            nl = data.find(b"\n", p)
            if nl < 0 or nl + 1 >= n:
                return None  # ID line without newline / "No sequence information given!"
            m = _NEXT_ID.search(data, nl + 1)
            e = m.start() if m else n
            seq = data[nl + 1 : e].translate(None, _WS_DIGITS)  # blanks and digits inside the sequence are skipped
            if seq.translate(None, _LEGAL):
                return None  # "Encountered an unexpected letter"
            out.append(seq)
            p = e
    else:
        while p < n:
            if data[p : p + 1] != b"@":
                return None  # "Expected '@' on beginning of ID line"
            nl = data.find(b"\n", p)
            if nl < 0:
                return None
            e = data.find(b"+", nl + 1)  # letters up to the first '+', blanks skipped
            if e < 0:
                return None
            seq = data[nl + 1 : e].translate(None, _WS)
            if seq.translate(None, _LEGAL):
                return None
            nl3 = data.find(b"\n", e)
            if nl3 < 0:
                return None
            q, need = nl3 + 1, len(seq)
            if q + need <= n and data.find(b"\n", q, q + need) < 0:
                q, need = q + need, 0
            while need and q < n:  # qualities: `need` characters that are not blanks, over as many lines as it takes
                end = data.find(b"\n", q)
                end = n if end < 0 else end + 1
                line = data[q:end]
                ns = len(line.translate(None, _WS))
                if ns < need:
                    need -= ns
                    q = end
                    continue
                for idx in range(len(line)):
                    if line[idx] not in _WS:
                        need -= 1
                        if need == 0:
                            q += idx + 1
                            break
            if need:
                return None  # "File ended before expected number of qualities could be read."
            if q < n:
                if data[q : q + 1] != b"\n":
                    return None  # "Qualitites longer than sequence."
                q += 1
            out.append(seq)
            p = q
    return out


class GpuBackend:
    """The device side of the build: K2 over the sequences of a target, filter creation and insertion in HBM."""

    def __init__(self, device: int = 0):
        from . import classify as _c

        self._c = _c
        self.device = device
        self.db = None

    def minimisers(self, seqs: Sequence[bytes], k: int, w: int):
        import numpy as np

        if not seqs:
            return np.empty(0, dtype=np.uint64)
        # K2 applies ganon-classify's rule "shorter than the window: skipped" (GC.cpp:690); the builder hashes every sequence
        # with seqan3's view, which shrinks the window to a sequence shorter than it (minimiser.hpp:298-299; the reference
        # builder counts 1 minimiser for 25 bp at k=19, w=31).  Those rare sequences go through the single-sequence hook,
        # which clamps the window; both run on the device.
        full = [s for s in seqs if len(s) >= w]
        parts = [self._c.minimisers(s, k, w, device=self.device) for s in seqs if k <= len(s) < w]
        if full:
            parts.append(self._c.minimisers_batch(full, k, w, device=self.device)[1])
        return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint64)

    def file_hashes(self, path: str, k: int, w: int, min_length: int):
        """count_hashes of one file inside the library: native reader (plain / gzip), K2 in segments, sort + unique in
        HBM.  Returns (distinct hashes or None on a parse error, n_sequences, n_skipped, n_bases)."""
        h, st = self._c.build_file_hashes(path, k, w, min_length, device=self.device, io_threads=2)
        return h, int(st.n_sequences), int(st.n_skipped), int(st.n_bases)

    def create(self, n_bins: int, bin_size_bits: int, hash_functions: int, k: int, w: int) -> None:
        self.db = self._c.Database.create(n_bins, bin_size_bits, hash_functions, k, w, device=self.device)

    def save(self, path: str, counts, layout, params) -> None:
        """The .ibf written by the library straight from HBM (no host copy of the bitvector)."""
        import numpy as np

        names = list(counts)
        index = {t: i for i, t in enumerate(names)}
        self.db.set_targets(names, np.array([index[t] for t, _f, _l in layout], dtype=np.uint32), np.array([counts[t] for t in names], dtype=np.uint64), params.max_hashes_bin)
        self.db.set_fp(params.max_fp, params.true_max_fp, params.true_avg_fp)
        self.db.save(path)

    def emplace(self, hashes, bins) -> None:
        self.db.emplace(hashes, bins)

    def words(self):
        i = self.db.info()
        return self.db.read_words(0, i.bin_size_bits * i.bin_words)

    def close(self) -> None:
        if self.db is not None:
            self.db.close()


def run_build(cfg: GanonBuildConfig, backend=None) -> bool:
    """GanonBuild::run.  Deterministic where the reference is not: targets are laid out in order of first appearance in
    the input table and a target's distinct minimisers in increasing order, so equal inputs give equal files."""
    import numpy as np

    from . import formats

    if not cfg.validate():
        return False
    t0 = time.time()
    targets, invalid = parse_input_table(cfg.input_file, cfg.quiet)
    if not targets:
        print("No valid input files", file=sys.stderr)
        return False
    own = backend is None
    be = backend or GpuBackend(cfg.device)
    try:
        # count_hashes (GanonBuild.cpp:184-249): distinct minimisers per target over all its files and sequences.  Like the
        # reference, every file's set is appended to <tmp-output-folder>/<target>.min when a folder is given, so that only
        # the counts stay in memory (RefSeq-scale builds); without a folder the sets are kept in host memory.
        spill = cfg.tmp_output_folder or None
        if spill:
            os.makedirs(spill, exist_ok=True)
        hashes: Dict[str, "np.ndarray"] = {}
        counts: Dict[str, int] = {}
        min_files: Dict[str, str] = {}
        n_seq = n_skipped = n_bp = 0
        native = hasattr(be, "file_hashes")
        results = None
        pool = None
        if native:
            # files are independent: several host threads read and index them while the device hashes others (the library
            # keeps per-call buffers and streams); results are consumed in input order, at most 64 files ahead
            from concurrent.futures import ThreadPoolExecutor

            todo = [f for files in targets.values() for f in files]
            pool = ThreadPoolExecutor(max(1, min(8, cfg.threads if cfg.threads > 1 else (os.cpu_count() or 1))))

            def windowed():
                for a in range(0, len(todo), 64):
                    yield from pool.map(lambda p: be.file_hashes(p, cfg.kmer_size, cfg.window_size, cfg.min_length), todo[a : a + 64])

            results = windowed()
        for target, files in targets.items():
            parts = []
            for f in files:
                if native:
                    u, ns, nk, nb = next(results)
                    n_seq += ns
                    n_skipped += nk
                    n_bp += nb
                    if u is None:
                        print("Error parsing file [%s]." % f, file=sys.stderr)
                        continue
                else:
                    seqs = []
                    records = read_sequences(f)
                    if records is None:
                        print("Error parsing file [%s]." % f, file=sys.stderr)
                        continue
                    for s in records:
                        if len(s) < cfg.min_length:
                            n_skipped += 1
                            continue
                        n_seq += 1
                        n_bp += len(s)
                        seqs.append(s)
                    u = np.unique(be.minimisers(seqs, cfg.kmer_size, cfg.window_size))
                parts.append(u)
            # the reference counts per file and adds up (a hash shared by two files of a target counts twice) and appends
            # every file's set to the target's .min file: keep the per-file sets concatenated
            counts[target] = int(sum(p.size for p in parts))
            if spill:
                min_files[target] = os.path.join(spill, target.replace("/", "_") + ".min")
                with open(min_files[target], "wb") as fh:
                    for p in parts:
                        p.tofile(fh)
            else:
                hashes[target] = np.concatenate(parts) if parts else np.empty(0, dtype=np.uint64)
        if pool is not None:
            pool.shutdown()
        params = choose_ibf_params(counts, cfg.max_fp, cfg.filter_size, cfg.hash_functions, cfg.mode)
        if params.n_bins == 0:
            print("No valid sequences to build", file=sys.stderr)
            return False
        layout = bin_layout(counts, params)
        be.create(params.n_bins, params.bin_size_bits, params.hash_functions, cfg.kmer_size, cfg.window_size)
        # insertion target by target, in bounded pieces (at most 2^25 hashes per emplace call)
        piece = 1 << 25
        by_target: Dict[str, List[Tuple[int, int, int]]] = {}
        for binno, (target, first, last) in enumerate(layout):
            by_target.setdefault(target, []).append((binno, first, last))
        for target, bins in by_target.items():
            th = np.fromfile(min_files[target], dtype=np.uint64) if spill else hashes[target]
            hs, bs, pending = [], [], 0
            for binno, first, last in bins:
                for a in range(first, last + 1, piece):
                    b = min(last + 1, a + piece)
                    hs.append(th[a:b])
                    bs.append(np.full(b - a, binno, dtype=np.uint32))
                    pending += b - a
                    if pending >= piece:
                        be.emplace(np.concatenate(hs), np.concatenate(bs))
                        hs, bs, pending = [], [], 0
            if pending:
                be.emplace(np.concatenate(hs), np.concatenate(bs))
            if spill:
                os.remove(min_files[target])
        if hasattr(be, "save"):
            be.save(cfg.output_file, counts, layout, params)
        else:
            db = formats.IBFFile(formats.IBF(params.n_bins, params.bin_size_bits, params.hash_functions, be.words()), cfg.kmer_size, cfg.window_size,
                                 params.max_hashes_bin, [(t, c) for t, c in counts.items()], [(b, t) for b, (t, _f, _l) in enumerate(layout)],
                                 max_fp=params.max_fp, true_max_fp=params.true_max_fp, true_avg_fp=params.true_avg_fp)
            formats.write_ibf(cfg.output_file, db)
    finally:
        if own:
            be.close()
    if not cfg.quiet:
        e = sys.stderr
        print("ganon-build processed %d sequences / %d files (%g Mbp) in %g seconds" % (n_seq, sum(len(f) for f in targets.values()), n_bp / 1e6, time.time() - t0), file=e)
        print(" - max. false positive: %g (avg.: %g)" % (params.true_max_fp, params.true_avg_fp), file=e)
        print(" - filter size: %gMB" % (params.filter_bits / float(MIB_BITS)), file=e)
        print(" - bins assigned: %d / hash functions: %d / max. hashes per bin: %d" % (params.n_bins, params.hash_functions, params.max_hashes_bin), file=e)
        if n_skipped or invalid:
            print(" - %d invalid files skipped, %d sequences shorter than --min-length skipped" % (invalid, n_skipped), file=e)
    return True


_BUILD_OPTS = {
    "input-file": ("i", str, "input_file"), "output-file": ("o", str, "output_file"), "kmer-size": ("k", int, "kmer_size"), "window-size": ("w", int, "window_size"),
    "hash-functions": ("s", int, "hash_functions"), "max-fp": ("p", float, "max_fp"), "filter-size": ("f", float, "filter_size"), "mode": ("j", str, "mode"),
    "min-length": ("y", int, "min_length"), "tmp-output-folder": ("m", str, "tmp_output_folder"), "threads": ("t", int, "threads"), "device": (None, int, "device"),
    "verbose": (None, bool, "verbose"), "quiet": (None, bool, "quiet"),
}


def build_main(argv: Optional[List[str]] = None) -> int:
    """`ganon-build` command line (src/ganon-build/CommandLineParser.cpp:15-85; exit codes of main.cpp)."""
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print("Try 'ganon-build -h/--help' for more information.", file=sys.stderr)
        return 1
    cfg = GanonBuildConfig()
    short = {v[0]: k for k, v in _BUILD_OPTS.items() if v[0]}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-h", "--help"):
            print("ganon-build (B200): " + ", ".join("--" + k for k in _BUILD_OPTS), file=sys.stderr)
            return 0
        if a in ("-v", "--version"):
            from .classify import VERSION

            print("version: " + VERSION, file=sys.stderr)
            return 0
        name, inline = (a[2:].partition("=")[0], a[2:].partition("=")[2] if "=" in a else None) if a.startswith("--") else (short.get(a[1:2]), a[2:] or None) if a.startswith("-") else (None, None)
        if name not in _BUILD_OPTS:
            print("Option '%s' does not exist" % a, file=sys.stderr)
            return 1
        _s, kind, attr = _BUILD_OPTS[name]
        if kind is bool:
            setattr(cfg, attr, True)
        else:
            if inline is None:
                i += 1
                if i >= len(argv):
                    print("Option '%s' is missing an argument" % name, file=sys.stderr)
                    return 1
                inline = argv[i]
            try:
                setattr(cfg, attr, kind(inline))
            except ValueError:
                print("Argument '%s' failed to parse" % inline, file=sys.stderr)
                return 1
        i += 1
    from ._lib import GnbError

    try:
        return 0 if run_build(cfg) else 1
    except GnbError as e:  # no device, out of memory ...: there is no CPU fallback
        print("ERROR: " + e.msg, file=sys.stderr)
        return 1
