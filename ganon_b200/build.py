"""Host arithmetic of `ganon-build`: the choice of the IBF parameters from the per-target minimiser
counts.  This is the part of the build that is not a kernel -- K2 (minimisers) and `k_emplace` (insertion) already exist
in the library; a GPU `ganon-build` is this module in front of them.

Mirrors `optimal_hashes` and its helpers (src/ganon-build/GanonBuild.cpp:290-618) and `true_false_positive` (382-412).
All quantities are IEEE doubles evaluated with libm in the reference's order, so the chosen configuration is the
reference's (pinned against the reference binary by tests/test_build_config_cpu.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

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


def _u64(x: float) -> int:
    """C++ conversion of a non-negative double to uint64_t (truncation)."""
    return int(x)


def bin_bits_for_fp(max_fp: float, n_hashes: int, hash_functions: int = 0) -> int:
    """Bits of one bin for `n_hashes` elements at false-positive rate `max_fp` (GanonBuild.cpp:290-305): with the optimal
    number of hash functions when none is given, else for that number."""
    if hash_functions == 0:
        return _u64(math.ceil((n_hashes * math.log(max_fp)) / math.log(1.0 / math.pow(2, math.log(2)))))
    return _u64(math.ceil(n_hashes * (-hash_functions / math.log(1 - math.exp(math.log(max_fp) / hash_functions)))))


def pick_hash_functions(bin_size_bits: int, n_hashes: int, requested: int, limit: int = 5) -> int:
    """Number of hash functions (GanonBuild.cpp:308-333): the requested one, or ln2 * bits/elements truncated to uint8,
    clamped to 1..limit (0 and values above the limit become the limit)."""
    h = requested
    if h == 0:
        h = int(math.log(2) * (bin_size_bits / float(n_hashes))) & 0xFF  # static_cast<uint8_t>
    if h > limit or h == 0:
        h = limit
    return h


def split_bins(counts: Iterable[int], n_hashes: int) -> int:
    """Bins needed when no bin holds more than n_hashes elements (GanonBuild.cpp:336-347)."""
    return sum(_u64(math.ceil(c / float(n_hashes))) for c in counts)


def padded_bins(n_bins: int) -> int:
    """Technical bins: the next multiple of 64 (GanonBuild.cpp:365-371)."""
    return _u64(math.ceil(n_bins / 64.0) * 64)


def bloom_fp(bin_size_bits: int, hash_functions: int, n_hashes: int) -> float:
    """Theoretical false-positive rate of one bin (GanonBuild.cpp:373-380)."""
    return math.pow(1 - math.exp(-hash_functions / (bin_size_bits / float(n_hashes))), hash_functions)


def split_correction(max_split_bins: int, max_fp: float, hash_functions: int, n_hashes: int) -> float:
    """Growth of a bin that keeps the rate of a target spread over `max_split_bins` bins at max_fp
    (multiple testing; GanonBuild.cpp:350-362)."""
    target_fpr = 1.0 - math.exp(math.log(1.0 - max_fp) / max_split_bins)
    grown = bin_bits_for_fp(target_fpr, n_hashes, hash_functions)
    base = bin_bits_for_fp(max_fp, n_hashes, hash_functions)
    return float(grown) / base  # ZeroDivisionError where the reference divides by an integer zero -> inf: handled by the caller


def real_fp(counts: Iterable[int], max_hashes_bin: int, bin_size_bits: int, hash_functions: int) -> Tuple[float, float]:
    """(highest, average) false-positive rate over the targets with their split bins (GanonBuild.cpp:382-412).  The
    average is summed in the order given (the reference sums in hash-map order: equal up to the last bits)."""
    highest = total = 0.0
    n = 0
    for c in counts:
        n_bins_target = _u64(math.ceil(c / float(max_hashes_bin)))
        n_hashes_bin = _u64(math.ceil(c / float(n_bins_target)))
        fp = 1.0 - math.pow(1.0 - bloom_fp(bin_size_bits, hash_functions, n_hashes_bin), n_bins_target)
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
    sims: List[Tuple[int, int, int, float]] = []  # (n_hashes, n_bins, filter_size_bits, fp)
    min_filter = min_bins = 0
    min_fp = 1.0
    step = min(100, largest)
    n = largest + 1
    while n > step:
        cap = n - 1
        bins = split_bins(counts, cap)
        if filter_size:
            bits = _u64((filter_size / float(padded_bins(bins))) * MIB_BITS)
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
            fp = 1 - math.pow(1.0 - bloom_fp(bits, h, cap), most_splits)
            min_fp = min(min_fp, fp)
        else:
            per_split = _u64(math.ceil(largest / float(most_splits)))
            approx = min(bloom_fp(bits, h, per_split), max_fp)
            try:
                rate = split_correction(most_splits, approx, h, cap)
            except ZeroDivisionError:
                rate = math.inf
            if math.isinf(rate) or math.isnan(rate):
                break
            bits = _u64(bits * rate)
            filter_bits = bits * padded_bins(bins)
            if filter_bits == 0:
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
        var_ratio = fp / min_fp if filter_size else filter_bits / float(min_filter)
        bins_ratio = bins / float(min_bins)
        score = (1 + math.pow(tilt, 2)) * ((var_ratio * bins_ratio) / ((w_var * var_ratio) + (w_bins * bins_ratio)))
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
        per_bin = min(_u64(math.ceil(count / float(n_bins_target))), params.max_hashes_bin)
        for i in range(n_bins_target):
            first = i * per_bin
            if first >= count:
                break
            out.append((target, first, min(first + per_bin, count) - 1))
    return out
