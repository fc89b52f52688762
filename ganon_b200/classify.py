"""Python host of the classify path.

Mirrors the two reference layers either side of the process boundary it replaces:

  * ``classify(cfg)``          <->  src/ganon/classify.py:7-107 (picks .hibf/.ibf/.tax per --db-prefix, assembles the
                                    ganon-classify arguments; here the call is in-process through the C ABI instead of
                                    ``subprocess``)
  * ``GanonClassifyConfig`` /
    ``run(config)``            <->  GanonClassify::Config (Config.hpp:22-49, validate 71-245) and GanonClassify::run
                                    (GC.cpp:1676): opens the output files, streams the read files block by block into
                                    ``gnb_session_classify`` and writes the text it returns.

All compute is in libganon_b200.so (CUDA); this module only moves bytes between files and the library.
"""
from __future__ import annotations

import ctypes as C
import gzip
import zlib
import os
import sys
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib
from ._lib import BatchResult, DbInfo, SessionConfig, Totals, check

VERSION = "ganon-b200 0.1.0 (ganon-classify 2.4.1 compatible)"
# bytes of a read file per batch.  64 MiB (~200 k reads of 150 bp): measured on B200 with the 8 GiB workload, 33.5 M reads
# from a 10.6 GB FASTQ take 0.73 s with 64 MiB blocks and 1.17 s with 256 MiB ones (page-locking the ring of buffers is
# a fixed cost of ~0.5 ms per MiB, and the parallel file reads hide behind the staging only with smaller blocks)
BLOCK_BYTES = int(os.environ.get("GANON_B200_BLOCK_BYTES", str(64 << 20)))


# ----------------------------------------------------------------------------------------------------------------------
# thin object wrappers over the C handles
# ----------------------------------------------------------------------------------------------------------------------
class Database:
    """A .ibf / .hibf resident in HBM (gnb_db)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def open(cls, path: str, hibf: bool = False, device: int = 0, shard: int = 0, n_shards: int = 1, hbm_budget: int = 0) -> "Database":
        """hbm_budget (bytes): a flat filter larger than this is loaded paged (host-resident tier, gnb_db_open_paged)."""
        h = C.c_void_p()
        if hbm_budget and not hibf and n_shards == 1:
            check(_lib.lib().gnb_db_open_paged(path.encode(), device, hbm_budget, C.byref(h)))
        else:
            check(_lib.lib().gnb_db_open(path.encode(), int(hibf), device, shard, n_shards, C.byref(h)))
        return cls(h.value)

    def page_out(self, hbm_budget: int) -> None:
        """Keep at most hbm_budget bytes of this (whole-in-HBM) filter on the device, the rest in page-locked host memory."""
        check(_lib.lib().gnb_db_page_out(self._h, hbm_budget))

    @classmethod
    def create(cls, bins: int, bin_size_bits: int, hash_functions: int, kmer_size: int, window_size: int, device: int = 0, shard: int = 0, n_shards: int = 1) -> "Database":
        h = C.c_void_p()
        check(_lib.lib().gnb_db_create_sharded(bins, bin_size_bits, hash_functions, kmer_size, window_size, device, shard, n_shards, C.byref(h)))
        return cls(h.value)

    @classmethod
    def create_hibf(cls, bins: Sequence[int], bin_size_bits: Sequence[int], hash_functions: int, kmer_size: int, window_size: int, next_ibf: Sequence[Sequence[int]],
                    bin_to_user: Sequence[Sequence[int]], user_bin_names: Sequence[str], fpr: float = 0.05, device: int = 0) -> "Database":
        """An HIBF in HBM: per sub-IBF the technical bins, rows, next_ibf_id and ibf_bin_to_filename_position rows (raptor layout)."""
        import numpy as np

        b = np.ascontiguousarray(bins, dtype=np.uint64)
        s = np.ascontiguousarray(bin_size_bits, dtype=np.uint64)
        nx = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.int64) for v in next_ibf]), dtype=np.int64)
        bu = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.int64) for v in bin_to_user]), dtype=np.int64)
        assert nx.size == bu.size == int(b.sum())
        names = (C.c_char_p * len(user_bin_names))(*[n.encode() for n in user_bin_names])
        h = C.c_void_p()
        check(_lib.lib().gnb_db_create_hibf(b.size, b.ctypes.data, s.ctypes.data, hash_functions, kmer_size, window_size, nx.ctypes.data, bu.ctypes.data, len(user_bin_names), names, fpr, device, C.byref(h)))
        return cls(h.value)

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def info(self) -> DbInfo:
        i = DbInfo()
        check(_lib.lib().gnb_db_info(self._h, C.byref(i)))
        return i

    def targets(self) -> List[Tuple[str, float, int]]:
        out = []
        name, fpr, nb = C.c_char_p(), C.c_double(), C.c_uint64()
        for i in range(self.info().n_targets):
            check(_lib.lib().gnb_db_target(self._h, i, C.byref(name), C.byref(fpr), C.byref(nb)))
            out.append((name.value.decode(), fpr.value, nb.value))
        return out

    def fill_random(self, seed: int, and_terms: int = 1) -> None:
        check(_lib.lib().gnb_db_fill_random(self._h, seed, and_terms))

    def emplace(self, hashes, bins, ibf_index: int = 0) -> None:
        import numpy as np

        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        bins = np.ascontiguousarray(bins, dtype=np.uint32)
        assert hashes.size == bins.size
        check(_lib.lib().gnb_db_emplace_ibf(self._h, ibf_index, hashes.ctypes.data, bins.ctypes.data, hashes.size))

    def set_targets(self, names: Sequence[str], bin_target, target_hashes, max_hashes_bin: int) -> None:
        import numpy as np

        bt = np.ascontiguousarray(bin_target, dtype=np.uint32)
        th = np.ascontiguousarray(target_hashes, dtype=np.uint64)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        check(_lib.lib().gnb_db_set_targets(self._h, len(names), arr, bt.ctypes.data, th.ctypes.data, max_hashes_bin))

    def set_fp(self, max_fp: float, true_max_fp: float, true_avg_fp: float) -> None:
        check(_lib.lib().gnb_db_set_fp(self._h, max_fp, true_max_fp, true_avg_fp))

    def read_words(self, offset: int, n: int, ibf_index: int = 0):
        import numpy as np

        out = np.empty(n, dtype=np.uint64)
        check(_lib.lib().gnb_db_read_words(self._h, ibf_index, offset, n, out.ctypes.data))
        return out

    def save(self, path: str) -> None:
        check(_lib.lib().gnb_db_save(self._h, path.encode()))

    def bulk_count(self, hashes, hash_off, ibf_index: int = 0):
        """counting_agent::bulk_count for several hash lists (test hook): uint16[n_reads, technical_bins]."""
        import numpy as np

        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        hash_off = np.ascontiguousarray(hash_off, dtype=np.uint64)
        n = hash_off.size - 1
        i = self.info()
        width = (i.shard_word_end - i.shard_word_begin) * 64 if ibf_index == 0 else None
        if width is None:
            raise NotImplementedError
        out = np.zeros((n, width), dtype=np.uint16)
        check(_lib.lib().gnb_db_bulk_count(self._h, ibf_index, hashes.ctypes.data, hash_off.ctypes.data, n, out.ctypes.data))
        return out

    def close(self) -> None:
        if self._h:
            _lib.lib().gnb_db_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_file_hashes(path: str, k: int, w: int, min_length: int = 0, device: int = 0, io_threads: int = 0):
    """ganon-build's count_hashes for one file on the device (gnb_build_file_hashes): (distinct minimisers ascending --
    None if the file has a parse error --, stats)."""
    import numpy as np

    h = C.c_void_p()
    st = _lib.BuildFileStats()
    check(_lib.lib().gnb_build_file_hashes(device, path.encode(), k, w, min_length, io_threads, C.byref(h), C.byref(st)))
    try:
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().gnb_hash_set_data(h, C.byref(p), C.byref(n)))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n.value,)).copy() if n.value else np.empty(0, dtype=np.uint64)
    finally:
        _lib.lib().gnb_hash_set_free(h)
    return (None if st.parse_error else arr), st


def minimisers_batch(seqs: Sequence[bytes], k: int, w: int, device: int = 0):
    """Minimiser hashes of many sequences (reads shorter than w give none): (hash_off uint64[n+1], hashes uint64[])."""
    import numpy as np

    blob = b"".join(seqs)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in seqs])
    hoff = np.zeros(len(seqs) + 1, dtype=np.uint64)
    L = _lib.lib()
    check(L.gnb_minimisers_batch(device, k, w, blob, off.ctypes.data, len(seqs), hoff.ctypes.data, None, 0))
    hashes = np.empty(int(hoff[-1]) + 1, dtype=np.uint64)
    check(L.gnb_minimisers_batch(device, k, w, blob, off.ctypes.data, len(seqs), hoff.ctypes.data, hashes.ctypes.data, hashes.size))
    return hoff, hashes[: int(hoff[-1])]


def minimisers(seq: bytes, k: int, w: int, device: int = 0):
    """seqan3::views::minimiser_hash of one sequence on the GPU (test hook for kernel K2)."""
    import numpy as np

    if isinstance(seq, str):
        seq = seq.encode()
    out = np.empty(max(len(seq), 1), dtype=np.uint64)
    n = C.c_uint64()
    check(_lib.lib().gnb_minimisers(device, k, w, seq, len(seq), out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].copy()


class Comm:
    """This process's place in a bin-sharded multi-GPU run (gnb_comm): one process per GPU, NCCL inside the library.
    `unique_id()` on one rank, the bytes carried to the others by the caller, then `Comm(id, rank, n_ranks, device)` on
    every rank (collective)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.GNB_COMM_ID_BYTES)
        check(_lib.lib().gnb_comm_unique_id(buf, _lib.GNB_COMM_ID_BYTES))
        return buf.raw

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int):
        assert len(unique_id) == _lib.GNB_COMM_ID_BYTES
        h = C.c_void_p()
        check(_lib.lib().gnb_comm_create(C.c_char_p(unique_id), rank, n_ranks, device, C.byref(h)))
        self._h = h
        self.rank, self.n_ranks, self.device = rank, n_ranks, device

    @property
    def handle(self):
        return self._h

    def nccl_version(self) -> int:
        v = C.c_int()
        check(_lib.lib().gnb_comm_info(self._h, None, None, None, C.byref(v)))
        return v.value

    def close(self) -> None:
        if self._h:
            _lib.lib().gnb_comm_free(self._h)
            self._h = C.c_void_p()


class Session:
    """One classification run over all hierarchy levels (gnb_session)."""

    def __init__(
        self,
        dbs: Sequence[Database],
        rel_cutoff: Sequence[float],
        rel_filter: Sequence[float],
        fpr_query: Sequence[float],
        hierarchy_labels: Optional[Sequence[str]] = None,
        tax_files: Optional[Sequence[str]] = None,
        skip_lca: bool = False,
        tax_root_node: str = "1",
        output_lca: bool = False,
        output_all: bool = False,
        output_unclassified: bool = False,
        output_single: bool = False,
        device: int = 0,
        host_threads: int = 0,
        n_reads: int = 400,
        quiet: bool = True,
        cuda_stream: int = 0,
        comm: Optional["Comm"] = None,
        sliced_ingest: bool = False,
    ):
        n = len(dbs)
        labels = list(hierarchy_labels) if hierarchy_labels else ["H1"] * n
        self._keep = dict(
            dbs=(C.c_void_p * n)(*[d.handle for d in dbs]),
            labels=(C.c_char_p * n)(*[s.encode() for s in labels]),
            cutoff=(C.c_double * n)(*rel_cutoff),
            tax=(C.c_char_p * n)(*[t.encode() for t in tax_files]) if tax_files else None,
            rel_filter=(C.c_double * len(rel_filter))(*rel_filter),
            fpr_query=(C.c_double * len(fpr_query))(*fpr_query),
            root=tax_root_node.encode(),
            db_objs=list(dbs),
        )
        k = self._keep
        cfg = SessionConfig(
            n,
            k["dbs"],
            k["labels"],
            k["cutoff"],
            k["tax"] if k["tax"] is not None else None,
            len(rel_filter),
            k["rel_filter"],
            k["fpr_query"],
            int(skip_lca),
            k["root"],
            int(output_lca),
            int(output_all),
            int(output_unclassified),
            int(output_single),
            device,
            host_threads,
            n_reads,
            int(quiet),
            C.c_void_p(cuda_stream) if cuda_stream else None,
            comm.handle if comm is not None else None,
            int(sliced_ingest),
        )
        self._keep["comm"] = comm
        h = C.c_void_p()
        check(_lib.lib().gnb_session_create(C.byref(cfg), C.byref(h)))
        self._h = h
        n_levels = C.c_uint32()
        check(_lib.lib().gnb_session_level_count(self._h, C.byref(n_levels)))
        self.level_labels = []
        lab = C.c_char_p()
        for i in range(n_levels.value):
            check(_lib.lib().gnb_session_level_label(self._h, i, C.byref(lab)))
            self.level_labels.append(lab.value.decode())
        self._filters_per_level = [sum(1 for x in labels if x == l) for l in self.level_labels]

    @staticmethod
    def _ptr(buf, n=None):
        """(address, length) of bytes / bytearray / memoryview / numpy / (addr, len) tuples without copying."""
        if buf is None:
            return None, 0
        if isinstance(buf, tuple):
            return C.c_void_p(buf[0]), buf[1]
        if isinstance(buf, bytes):
            return C.cast(C.c_char_p(buf), C.c_void_p), len(buf) if n is None else n
        if isinstance(buf, bytearray):
            arr = (C.c_char * len(buf)).from_buffer(buf)
            return C.cast(arr, C.c_void_p), len(buf) if n is None else n
        if hasattr(buf, "ctypes"):  # numpy
            return C.c_void_p(buf.ctypes.data), buf.nbytes if n is None else n
        if hasattr(buf, "data_ptr"):  # torch (host tensor)
            return C.c_void_p(buf.data_ptr()), buf.numel() * buf.element_size() if n is None else n
        raise TypeError("unsupported buffer type %r" % type(buf))

    def classify(self, block1, block2=None, final: bool = True, prefix_id: int = 0, len1: Optional[int] = None, len2: Optional[int] = None) -> BatchResult:
        p1, n1 = self._ptr(block1, len1)
        p2, n2 = self._ptr(block2, len2)
        res = BatchResult()
        check(_lib.lib().gnb_session_classify(self._h, prefix_id, p1, n1, p2, n2, int(final), C.byref(res)))
        return res

    def stage(self, block1, block2=None, final: bool = True, len1=None, len2=None) -> int:
        p1, n1 = self._ptr(block1, len1)
        p2, n2 = self._ptr(block2, len2)
        n = C.c_uint64()
        check(_lib.lib().gnb_session_stage(self._h, p1, n1, p2, n2, int(final), C.byref(n)))
        return n.value

    def run_staged(self) -> BatchResult:
        res = BatchResult()
        check(_lib.lib().gnb_session_run_staged(self._h, C.byref(res)))
        return res

    def finish_staged(self, prefix_id: int = 0) -> BatchResult:
        res = BatchResult()
        check(_lib.lib().gnb_session_finish_staged(self._h, prefix_id, C.byref(res)))
        return res

    def submit(self, block1, block2=None, final: bool = True, prefix_id: int = 0, len1: Optional[int] = None, len2: Optional[int] = None) -> BatchResult:
        """Asynchronous classify: returns the staging info (n_reads, consumed1/2, parse_error); see collect()."""
        p1, n1 = self._ptr(block1, len1)
        p2, n2 = self._ptr(block2, len2)
        info = BatchResult()
        check(_lib.lib().gnb_session_submit(self._h, prefix_id, p1, n1, p2, n2, int(final), C.byref(info)))
        return info

    def collect(self) -> BatchResult:
        res = BatchResult()
        check(_lib.lib().gnb_session_collect(self._h, C.byref(res)))
        return res

    def classify_files(self, prefix_id: int, file1: str, file2: Optional[str] = None, fds: Optional["_lib.OutputFds"] = None, block_bytes: int = 0, io_threads: int = 0) -> "_lib.FilesResult":
        """One read file (or pair) start to end inside the library (gnb_session_classify_files); text goes to `fds`."""
        res = _lib.FilesResult()
        check(_lib.lib().gnb_session_classify_files(self._h, prefix_id, file1.encode(), file2.encode() if file2 else None, C.byref(fds) if fds is not None else None, block_bytes, io_threads, C.byref(res)))
        return res

    def in_flight(self) -> Tuple[int, int]:
        n, cap = C.c_uint32(), C.c_uint32()
        check(_lib.lib().gnb_session_in_flight(self._h, C.byref(n), C.byref(cap)))
        return n.value, cap.value

    # level-wise form (bin-sharded multi-GPU, see ganon_b200/sharded.py)
    def run_level(self, level: int) -> None:
        check(_lib.lib().gnb_session_run_level(self._h, level))

    def level_tuples(self, level: int, filt: int):
        import numpy as np

        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().gnb_session_level_tuples(self._h, level, filt, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.empty(0, dtype=np.uint64)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n.value,)).copy()

    def set_level_tuples(self, level: int, filt: int, tuples) -> None:
        import numpy as np

        t = np.ascontiguousarray(tuples, dtype=np.uint64)
        check(_lib.lib().gnb_session_set_level_tuples(self._h, level, filt, t.ctypes.data, t.size))

    def finish_level(self, level: int) -> None:
        check(_lib.lib().gnb_session_finish_level(self._h, level))

    # device form of the level-wise API (tuples stay in HBM; see ganon_b200/sharded.py)
    def run_level_device(self, level: int) -> None:
        check(_lib.lib().gnb_session_run_level_device(self._h, level))

    def level_tuples_device(self, level: int) -> Tuple[int, int]:
        """(device pointer, count) of the level's sorted tuples."""
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().gnb_session_level_tuples_device(self._h, level, C.byref(p), C.byref(n)))
        return (p.value or 0), n.value

    def set_level_tuples_device(self, level: int, dev_ptr: int, n: int) -> None:
        check(_lib.lib().gnb_session_set_level_tuples_device(self._h, level, C.c_void_p(dev_ptr), n))

    def finish_level_device(self, level: int, prefix_id: int = 0) -> None:
        check(_lib.lib().gnb_session_finish_level_device(self._h, level, prefix_id))

    def hibf_rounds(self):
        """[(kernel ms, algorithmic bytes, items)] per traversal round of the last HIBF filter run (staged forms)."""
        import numpy as np

        ms, by, it, n = np.zeros(16, np.float32), np.zeros(16, np.uint64), np.zeros(16, np.uint64), C.c_uint32()
        check(_lib.lib().gnb_session_hibf_rounds(self._h, 16, ms.ctypes.data, by.ctypes.data, it.ctypes.data, C.byref(n)))
        return [(float(ms[i]), int(by[i]), int(it[i])) for i in range(min(16, n.value))]

    def staged_timings(self) -> BatchResult:
        res = BatchResult()
        check(_lib.lib().gnb_session_staged_timings(self._h, C.byref(res)))
        return res

    def collect_staged(self, prefix_id: int = 0) -> BatchResult:
        res = BatchResult()
        check(_lib.lib().gnb_session_collect_staged(self._h, prefix_id, C.byref(res)))
        return res

    @property
    def n_filters_per_level(self) -> List[int]:
        return self._filters_per_level

    def node_name(self, level: int, node: int) -> str:
        s = C.c_char_p()
        check(_lib.lib().gnb_session_node_name(self._h, level, node, C.byref(s)))
        return s.value.decode()

    def report(self, prefix_id: int = 0) -> bytes:
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().gnb_session_report(self._h, prefix_id, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def stats(self, prefix_id: int = 0, prefix_name: str = "") -> bytes:
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().gnb_session_stats(self._h, prefix_id, prefix_name.encode(), C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def keep_matches(self, enable: bool = True) -> None:
        """Keep the matches of every classified read in HBM for `reassign` (call before the first batch)."""
        check(_lib.lib().gnb_session_keep_matches(self._h, int(enable)))

    def reassign(self, prefix_id: int = 0, threshold: float = 0.0, max_iter: int = 10):
        """EM reassignment (src/ganon/reassign.py) on the kept matches: ({group label: `.one` bytes}, new `.rep` bytes,
        {group label: (iterations, reads with several matches)})."""
        r = _lib.ReassignResult()
        check(_lib.lib().gnb_session_reassign(self._h, prefix_id, float(threshold), int(max_iter), C.byref(r)))
        ones, info = {}, {}
        for g in range(r.n_groups):
            lab = r.group_label[g].decode()
            ones[lab] = C.string_at(r.one_text[g], r.one_len[g]) if r.one_len[g] else b""
            info[lab] = (r.iterations[g], r.reassigned_reads[g])
        return ones, (C.string_at(r.rep_text, r.rep_len) if r.rep_len else b""), info

    def totals(self, prefix_id: int = 0, level: int = -1) -> Totals:
        t = Totals()
        check(_lib.lib().gnb_session_totals(self._h, prefix_id, level, C.byref(t)))
        return t

    def close(self) -> None:
        if self._h:
            _lib.lib().gnb_session_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def result_text(res: BatchResult, kind: str, level: int = 0) -> bytes:
    if kind == "unc":
        return C.string_at(res.unc_text, res.unc_len) if res.unc_len else b""
    ptrs, lens = (res.all_text, res.all_len) if kind == "all" else (res.one_text, res.one_len)
    return C.string_at(ptrs[level], lens[level]) if lens[level] else b""


# ----------------------------------------------------------------------------------------------------------------------
# GanonClassify::Config + run
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class GanonClassifyConfig:
    """Field for field GanonClassify::Config (Config.hpp:22-49)."""

    single_reads: List[str] = field(default_factory=list)
    paired_reads: List[str] = field(default_factory=list)
    batch_reads: List[str] = field(default_factory=list)
    ibf: List[str] = field(default_factory=list)
    tax: List[str] = field(default_factory=list)
    output_prefix: str = ""
    hierarchy_labels: List[str] = field(default_factory=lambda: ["H1"])
    rel_cutoff: List[float] = field(default_factory=lambda: [0.2])
    rel_filter: List[float] = field(default_factory=lambda: [0.0])
    fpr_query: List[float] = field(default_factory=lambda: [1.0])
    output_lca: bool = False
    output_all: bool = False
    output_unclassified: bool = False
    output_stats: bool = False
    output_single: bool = False
    hibf: bool = False
    skip_lca: bool = False
    tax_root_node: str = "1"
    threads: int = 1
    n_batches: int = 1000
    n_reads: int = 400
    verbose: bool = False
    quiet: bool = False
    # not in the reference: which GPU to use, or several GPUs with every .ibf bin-sharded over them (one process per GPU)
    device: int = 0
    devices: List[int] = field(default_factory=list)
    hbm_budget_gb: float = 0.0  # > 0: flat filters above this size are loaded paged (host-resident tier)
    # not in the reference binary: the EM step of `ganon classify` (src/ganon/reassign.py) from the matches in HBM
    reassign_em: bool = False
    em_max_iter: int = 10
    em_threshold: List[float] = field(default_factory=lambda: [0.0])
    em_write_one: bool = True  # reassign.py --skip-one: only the new .rep is written

    def _err(self, msg: str) -> bool:
        print(msg, file=sys.stderr)
        return False

    def _check_files(self, files: Sequence[str]) -> bool:
        for f in files:
            if not os.path.exists(f):
                return self._err("file not found: " + f)
            if os.path.getsize(f) == 0:
                return self._err("file is empty: " + f)
        return True

    def validate(self) -> bool:
        """Config::validate (Config.hpp:71-173) and validate_hierarchy (175-245), same messages."""
        if not self.output_prefix:
            return self._err("--output-prefix is mandatory")
        if not self.paired_reads and not self.single_reads and not self.batch_reads:
            return self._err("At least one of --[single|paired|batch]-reads is mandatory")
        if not self.ibf:
            return self._err("--ibf is mandatory")
        if (self.paired_reads or self.single_reads) and self.batch_reads:
            return self._err("--batch-reads cannot be used together with --[single|paired]-reads")
        if len(self.paired_reads) % 2 != 0:
            return self._err("--paired-reads should be an even number of files (pairs)")
        for files in (self.single_reads, self.paired_reads, self.batch_reads, self.ibf, self.tax):
            if not self._check_files(files):
                return False
        if any(v < 0 or v > 1 for v in self.rel_cutoff):
            return self._err("--rel-cutoff values should be set between 0 and 1 (0 to disable)")
        if any(v < 0 or v > 1 for v in self.rel_filter):
            return self._err("--rel-filter values should be set between 0 and 1 (1 to disable)")
        if any(v < 0 or v > 1 for v in self.fpr_query):
            return self._err("--fpr-query values should be set between 0 and 1 (1 to disable)")
        self.n_batches = max(1, self.n_batches)
        self.n_reads = max(1, self.n_reads)
        # validate_hierarchy
        unique = len(set(self.hierarchy_labels))
        if len(self.rel_filter) == 1 and unique > 1:
            self.rel_filter = self.rel_filter * unique
        elif len(self.rel_filter) != unique:
            return self._err("Please provide a single or one-per-hierarchy --rel-filter value[s]")
        if len(self.fpr_query) == 1 and unique > 1:
            self.fpr_query = self.fpr_query * unique
        elif len(self.fpr_query) != unique:
            return self._err("Please provide a single or one-per-hierarchy --fpr-query value[s]")
        if self.tax and len(self.ibf) != len(self.tax):
            return self._err("The number of files provided with --ibf and --tax should match")
        if len(self.hierarchy_labels) == 1 and len(self.ibf) > 1:
            self.hierarchy_labels = self.hierarchy_labels * len(self.ibf)
        elif len(self.hierarchy_labels) != len(self.ibf):
            return self._err("--hierarchy does not match with the number of --ibf and --tax")
        if len(self.rel_cutoff) == 1 and len(self.ibf) > 1:
            self.rel_cutoff = self.rel_cutoff * len(self.ibf)
        elif len(self.rel_cutoff) != len(self.ibf):
            return self._err("Please provide a single or one-per-filter --rel-cutoff value[s]")
        if not self.tax:
            self.skip_lca = True
        return True


IO_THREADS = int(os.environ.get("GANON_B200_IO_THREADS", "0"))  # 0 = all host threads but two (at most 16; split between paired gzip files)


def _parse_reads_config(cfg: GanonClassifyConfig) -> Optional[Dict[str, List[Tuple[str, str]]]]:
    """parse_reads_config (GC.cpp:287-351)."""
    rc: Dict[str, List[Tuple[str, str]]] = {}
    if cfg.batch_reads:
        for bf in cfg.batch_reads:
            with open(bf) as fh:
                lines = fh.read().split("\n")
                if lines and lines[-1] == "":
                    lines.pop()  # std::getline: the text after the last newline is a line only if it is not empty
                for line in lines:
                    fields = line.split("\t")
                    if fields[-1] == "":
                        fields.pop()  # getline on the fields: nothing follows a trailing tab, and an empty line has no field
                    if len(fields) <= 1:  # this includes blank lines
                        print("ERROR: invalid --batch-reads file (prefix <tab> file1 [<tab> file2])", file=sys.stderr)
                        return None
                    for p in fields[1:3] if len(fields) == 3 else fields[1:2]:
                        if not os.path.exists(p) or os.path.getsize(p) == 0:
                            print("ERROR: file not found/empty: " + p, file=sys.stderr)
                            return None
                    # exactly three fields make a pair; with four or more only the first file is used
                    rc.setdefault(fields[0], []).append((fields[1], fields[2] if len(fields) == 3 else ""))
    else:
        for f in cfg.single_reads:
            rc.setdefault("", []).append((f, ""))
        for i in range(0, len(cfg.paired_reads), 2):
            rc.setdefault("", []).append((cfg.paired_reads[i], cfg.paired_reads[i + 1]))
    return dict(sorted(rc.items()))  # std::map order


def run(cfg: GanonClassifyConfig) -> bool:
    """GanonClassify::run (GC.cpp:1676-1691) -> ganon_classify<TFilter> (GC.cpp:1375-1674).  With several --devices the
    databases are bin-sharded: this process is rank 0 (it writes every output file), one helper process per further GPU
    runs the same loop on its column shards; the library exchanges the sparse matches between them (NCCL)."""
    if not cfg.validate():
        return False
    if len(set(cfg.devices)) != len(cfg.devices):
        print("--devices lists a device twice", file=sys.stderr)
        return False
    if len(cfg.devices) > 1:
        if cfg.hibf:
            print("--devices needs flat .ibf databases (an HIBF descends per read: use one GPU)", file=sys.stderr)
            return False
        return _run_sharded(cfg)
    if cfg.devices:
        cfg.device = cfg.devices[0]
    return _run_rank(cfg, 0, 1, None)


def _helper_rank(cfg: GanonClassifyConfig, rank: int, n_ranks: int, uid: bytes) -> None:
    cfg.quiet = True
    ok = _run_rank(cfg, rank, n_ranks, uid)
    sys.exit(0 if ok else 1)


def _run_sharded(cfg: GanonClassifyConfig) -> bool:
    import multiprocessing as mp
    import threading

    n = len(cfg.devices)
    uid = Comm.unique_id()
    ctx = mp.get_context("spawn")
    helpers = [ctx.Process(target=_helper_rank, args=(cfg, r, n, uid), daemon=True) for r in range(1, n)]
    for h in helpers:
        h.start()
    stop = threading.Event()

    def watchdog() -> None:  # a helper that dies leaves the other ranks waiting in a collective: end the run instead
        while not stop.wait(0.5):
            for r, h in enumerate(helpers, 1):
                if h.exitcode not in (None, 0):
                    print("ERROR: the process of device %d ended with exit code %d" % (cfg.devices[r], h.exitcode), file=sys.stderr)
                    os._exit(1)

    threading.Thread(target=watchdog, daemon=True).start()
    try:
        ok = _run_rank(cfg, 0, n, uid)
    finally:
        stop.set()
    for h in helpers:
        h.join(timeout=60 if ok else 1)
        if h.is_alive():
            h.terminate()
            ok = False
    return ok and all(h.exitcode == 0 for h in helpers)


def _run_rank(cfg: GanonClassifyConfig, rank: int, n_ranks: int, uid: Optional[bytes]) -> bool:
    """The run of one process on one GPU; rank 0 (the only rank of an unsharded run) owns the output files."""
    writer = rank == 0
    device = cfg.devices[rank] if n_ranks > 1 else cfg.device
    t_start = time.time()
    reads_config = _parse_reads_config(cfg)
    if reads_config is None:
        return False
    for prefix in reads_config:
        d = os.path.dirname(cfg.output_prefix + prefix)
        if d and not os.path.isdir(d):
            os.makedirs(d, exist_ok=True)

    # ---- load every database into HBM (load_files GC.cpp:1007-1039) ----
    t_load = time.time()
    dbs: List[Database] = []
    try:
        for path in cfg.ibf:
            dbs.append(Database.open(path, hibf=cfg.hibf, device=device, shard=rank, n_shards=n_ranks, hbm_budget=int(cfg.hbm_budget_gb * (1 << 30)) if n_ranks == 1 else 0))
        comm = Comm(uid, rank, n_ranks, device) if n_ranks > 1 else None
    except _lib.GnbError as e:
        print("ERROR: loading ibf or tax files (%s)" % e.msg, file=sys.stderr)
        return False
    t_load = time.time() - t_load
    try:
        sess = Session(
            dbs,
            cfg.rel_cutoff,
            cfg.rel_filter,
            cfg.fpr_query,
            hierarchy_labels=cfg.hierarchy_labels,
            tax_files=cfg.tax or None,
            skip_lca=cfg.skip_lca,
            tax_root_node=cfg.tax_root_node,
            output_lca=cfg.output_lca,
            output_all=cfg.output_all,
            output_unclassified=cfg.output_unclassified,
            output_single=cfg.output_single,
            device=device,
            host_threads=cfg.threads if cfg.threads > 1 else 0,
            n_reads=cfg.n_reads,
            quiet=cfg.quiet,
            comm=comm,
            sliced_ingest=comm is not None,
        )
    except _lib.GnbError as e:
        print(e.msg, file=sys.stderr)
        return False

    labels = sess.level_labels
    multi = len(labels) > 1 and not cfg.output_single
    write_one = cfg.output_lca and not cfg.skip_lca
    if cfg.reassign_em:
        if write_one:
            print("--reassign-em writes the .one file itself: it cannot be combined with --output-lca", file=sys.stderr)
            return False
        sess.keep_matches(True)
    prefixes = list(reads_config)
    # every rank of a sharded run holds the complete result; only rank 0 writes it
    open_out = (lambda path: open(path, "wb")) if writer else (lambda path: open(os.devnull, "wb"))
    out_rep = {p: open_out(cfg.output_prefix + p + ".rep") for p in prefixes}
    out_unc = {p: open_out(cfg.output_prefix + p + ".unc") for p in prefixes} if cfg.output_unclassified else {}

    def level_files(ext: str) -> Dict[str, List]:
        files: Dict[str, List] = {}
        for p in prefixes:
            if multi:
                files[p] = [open_out(cfg.output_prefix + p + "." + lab + "." + ext) for lab in labels]
            else:
                fh = open_out(cfg.output_prefix + p + "." + ext)
                files[p] = [fh] * len(labels)
        return files

    out_all = level_files("all") if cfg.output_all else {}
    out_one = level_files("one") if write_one else {}

    # reader -> classify -> writer (GC.cpp:1220-1322) runs inside the library per file (pair): block ring in page-locked
    # memory, parallel preads / parallel inflate, batches in flight on the GPU, a writer thread on the files opened here
    t_class = time.time()
    prof = {"open": 0.0, "read_wait": 0.0, "submit": 0.0, "collect": 0.0, "write": 0.0, "blocks": 0}
    n_lv = len(labels)
    try:
        for pid, prefix in enumerate(prefixes):
            for fh in set(out_all.get(prefix, []) + out_one.get(prefix, []) + ([out_unc[prefix]] if prefix in out_unc else [])):
                fh.flush()
            all_fd = (C.c_int * n_lv)(*[out_all[prefix][li].fileno() if cfg.output_all else -1 for li in range(n_lv)])
            one_fd = (C.c_int * n_lv)(*[out_one[prefix][li].fileno() if write_one else -1 for li in range(n_lv)])
            fds = _lib.OutputFds(n_lv, all_fd, one_fd, out_unc[prefix].fileno() if cfg.output_unclassified else -1)
            for file1, file2 in reads_config[prefix]:
                fr = sess.classify_files(pid, file1, file2 or None, fds if writer else None, BLOCK_BYTES, IO_THREADS)
                prof["open"] += fr.ms_open / 1e3
                prof["read_wait"] += fr.ms_read_wait / 1e3
                prof["submit"] += fr.ms_submit / 1e3
                prof["collect"] += fr.ms_collect / 1e3
                prof["write"] += fr.ms_write / 1e3
                prof["blocks"] += fr.n_blocks
    except _lib.GnbError as e:
        print("ERROR: %s" % e.msg, file=sys.stderr)
        for group in (list(out_unc.values()), *(v for v in out_all.values()), *(v for v in out_one.values()), list(out_rep.values())):
            for fh in group:
                if not fh.closed:
                    fh.close()
        sess.close()
        for d in dbs:
            d.close()
        return False
    t_class = time.time() - t_class

    for pid, prefix in enumerate(prefixes):
        if not writer:
            out_rep[prefix].close()
            continue
        if cfg.reassign_em:
            # reassign.py: `.one` (one per hierarchy label unless there is a single `.all`) and the new `.rep`
            ones, new_rep, info = sess.reassign(pid, cfg.em_threshold[0], cfg.em_max_iter)
            for lab, text in ones.items():
                if cfg.em_write_one:
                    with open(cfg.output_prefix + prefix + ("." + lab if lab else "") + ".one", "wb") as fh:
                        fh.write(text)
                if not cfg.quiet:
                    print(" - %d iteration(s), %d reassigned reads%s" % (info[lab][0], info[lab][1], " [" + lab + "]" if lab else ""), file=sys.stderr)
            out_rep[prefix].write(new_rep)
        else:
            out_rep[prefix].write(sess.report(pid))
        out_rep[prefix].close()
        if cfg.output_stats:
            with open(cfg.output_prefix + prefix + ".sta", "wb") as fh:
                fh.write(sess.stats(pid, prefix))
    seen = set()
    for group in (out_unc.values(), *(v for v in out_all.values()), *(v for v in out_one.values())):
        for fh in group if not isinstance(group, list) else group:
            if id(fh) not in seen:
                seen.add(id(fh))
                fh.close()

    if cfg.verbose and not cfg.quiet:
        print("host pipeline (s): open+pin %.3f, waiting for file blocks %.3f, staging (H2D + record index) %.3f, waiting for results %.3f, writer thread %.3f; %d blocks of <= %d MiB" % (prof["open"], prof["read_wait"], prof["submit"], prof["collect"], prof["write"], prof["blocks"], BLOCK_BYTES >> 20), file=sys.stderr)
    if not cfg.quiet and writer:
        _print_stats(cfg, sess, prefixes, labels, t_class, t_load, time.time() - t_start)
    sess.close()
    for d in dbs:
        d.close()
    if comm is not None:
        comm.close()
    return True


def _print_stats(cfg, sess: Session, prefixes, labels, t_class: float, t_load: float, t_total: float) -> None:
    """print_time / print_stats (GC.cpp:1041-1128), same wording."""
    e = sys.stderr
    if cfg.verbose:
        print("loading filter(s)    elapsed (s): %g seconds" % t_load, file=e)
        print("classifying+printing elapsed (s): %g seconds" % t_class, file=e)
        print("total                elapsed (s): %g seconds" % t_total, file=e)
        print("-" * 70 + "\n", file=e)
    tot = [sess.totals(i) for i in range(len(prefixes))]
    seqs = sum(t.seqs_processed for t in tot)
    length = sum(t.length_processed for t in tot)
    kmers = sum(t.kmers_processed for t in tot)
    print("ganon-classify processed %d sequences (%g Mbp) with %d k-mers in %g seconds (%g Mbp/m)" % (seqs, length / 1e6, kmers, t_class, (length / 1e6) / (max(t_class, 1e-9) / 60.0)), file=e)

    def db(t: Totals, seq_processed: float, seq_unclassified: int) -> None:
        multiple = t.seqs_classified - t.seqs_unique
        avg = t.matches / t.seqs_classified if t.seqs_classified else 0
        perc = t.kmers_matches / t.kmers_from_classified_seqs * 100 if t.kmers_matches else 0
        print("%d sequences classified (%g%%)" % (t.seqs_classified, t.seqs_classified / seq_processed * 100), file=e)
        print("  %d with unique matches (%g%%)" % (t.seqs_unique, t.seqs_unique / seq_processed * 100), file=e)
        print("  %d with multiple matches (%g%%)" % (multiple, multiple / seq_processed * 100), file=e)
        if seq_unclassified > 0:
            print("%d sequences unclassified (%g%%)" % (seq_unclassified, seq_unclassified / seq_processed * 100), file=e)
            if t.seqs_skipped_small:
                print("  %d sequences skipped (shorter than window size)" % t.seqs_skipped_small, file=e)
            if t.seqs_skipped_big:
                print("  %d sequences skipped (larger than allowed, check compilation with -DLONGREADS)" % t.seqs_skipped_big, file=e)
        print("matches: %d (avg. %g reference/sequence), %d discarded (--rel-filter), %d discarded (--fpr-query)" % (t.matches, avg, t.discarded_matches_filter, t.discarded_matches_fprquery), file=e)
        print("k-mers: %d/%d k-mers matched/k-mers from classified sequences (%g%%)" % (t.kmers_matches, t.kmers_from_classified_seqs, perc), file=e)

    for pid, prefix in enumerate(prefixes):
        t = tot[pid]
        if len(prefixes) > 1:
            print("\n[%s] %d sequences (%g Mbp) with %d k-mers" % (prefix, t.seqs_processed, t.length_processed / 1e6, t.kmers_processed), file=e)
        sp = float(t.seqs_processed) if t.seqs_processed > 0 else 1.0
        db(t, sp, t.seqs_processed - t.seqs_classified)
        if len(labels) > 1:
            print("\nBy database hierarchical level:", file=e)
            for li, lab in enumerate(labels):
                print(lab + ":", file=e)
                db(sess.totals(pid, li), sp, 0)


# ----------------------------------------------------------------------------------------------------------------------
# src/ganon/classify.py
# ----------------------------------------------------------------------------------------------------------------------
def classify(cfg) -> bool:
    """Drop-in for ``ganon.classify.classify(cfg)`` up to the ganon-classify call (src/ganon/classify.py:7-64): the
    same selection of .hibf/.ibf/.tax per --db-prefix and the same argument mapping; reassign/report stay the
    reference's (they consume the files written here).  With ``cfg.reassign_in_memory`` set and --multiple-matches em
    the EM step (src/ganon/classify.py:76-88 -> reassign.py) runs on the matches kept in HBM instead: `.one` is written if
    --output-one, `.all` only if --output-all, and the caller skips its own reassign() call."""
    filter_files, tax_files, hibf = [], [], False
    for db_prefix in cfg.db_prefix:
        if os.path.isfile(db_prefix + ".hibf") and os.path.getsize(db_prefix + ".hibf") > 0:
            filter_files.append(db_prefix + ".hibf")
            hibf = True
        elif os.path.isfile(db_prefix + ".ibf") and os.path.getsize(db_prefix + ".ibf") > 0:
            filter_files.append(db_prefix + ".ibf")
        if os.path.isfile(db_prefix + ".tax") and os.path.getsize(db_prefix + ".tax") > 0:
            tax_files.append(db_prefix + ".tax")
    if len(tax_files) != len(filter_files):
        tax_files = []
    g = lambda name, default=None: getattr(cfg, name, default)
    mm = g("multiple_matches", "em")
    c = GanonClassifyConfig(
        single_reads=list(g("single_reads") or []),
        paired_reads=list(g("paired_reads") or []),
        batch_reads=list(g("batch_reads") or []),
        ibf=filter_files,
        tax=tax_files,
        output_prefix=g("output_prefix", "") or "",
        hierarchy_labels=list(g("hierarchy_labels") or ["H1"]),
        # an empty value leaves the flag out and with it the binary's default (Config.hpp:36-38), classify.py:40-48
        rel_cutoff=[float(x) for x in (g("rel_cutoff") or [0.2])],
        rel_filter=[float(x) for x in (g("rel_filter") or [0.0])],
        fpr_query=[float(x) for x in (g("fpr_query") or [1.0])],
        skip_lca=mm != "lca",
        output_lca=mm == "lca" and bool(g("output_one")),
        output_all=bool(g("output_all")) or (mm == "em" and not g("reassign_in_memory")),
        reassign_em=mm == "em" and bool(g("reassign_in_memory")),
        em_write_one=bool(g("output_one")),
        em_max_iter=10 if g("max_iter") is None else int(g("max_iter")),  # 0 = until convergence
        em_threshold=[float(g("threshold", 0) or 0)],
        output_unclassified=bool(g("output_unclassified")),
        output_stats=bool(g("output_stats")),
        output_single=bool(g("output_single")),
        threads=int(g("threads", 1) or 1),
        verbose=bool(g("verbose")),
        hibf=bool(g("hibf", hibf)),
        quiet=bool(g("quiet")),
    )
    if g("n_reads") is not None:
        c.n_reads = int(g("n_reads"))
    if g("n_batches") is not None:
        c.n_batches = int(g("n_batches"))
    return run(c)
