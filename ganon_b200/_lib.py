"""ctypes binding of libganon_b200.so (include/ganon_b200.h).

The library is the product: if it is missing this module raises -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libganon_b200.so")
CSRC = os.path.join(_HERE, "csrc")

GNB_OK = 0
GNB_COMM_ID_BYTES = 256
STATUS = {0: "GNB_OK", -1: "GNB_ERR_ARG", -2: "GNB_ERR_IO", -3: "GNB_ERR_FORMAT", -4: "GNB_ERR_CUDA", -5: "GNB_ERR_CONFIG", -6: "GNB_ERR_PARSE", -7: "GNB_ERR_LIMIT"}


class GnbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (STATUS.get(code, str(code)), msg))
        self.code = code
        self.msg = msg


class DbInfo(C.Structure):
    _fields_ = [
        ("is_hibf", C.c_int),
        ("kmer_size", C.c_uint32),
        ("window_size", C.c_uint32),
        ("hash_functions", C.c_uint32),
        ("bins", C.c_uint64),
        ("technical_bins", C.c_uint64),
        ("bin_size_bits", C.c_uint64),
        ("bin_words", C.c_uint64),
        ("shard_word_begin", C.c_uint64),
        ("shard_word_end", C.c_uint64),
        ("max_hashes_bin", C.c_uint64),
        ("max_fp", C.c_double),
        ("n_targets", C.c_uint64),
        ("n_ibfs", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("device", C.c_int),
        ("n_pages", C.c_uint64),
        ("n_resident_pages", C.c_uint64),
        ("host_bytes", C.c_uint64),
    ]


class SessionConfig(C.Structure):
    _fields_ = [
        ("n_filters", C.c_uint32),
        ("dbs", C.POINTER(C.c_void_p)),
        ("hierarchy_labels", C.POINTER(C.c_char_p)),
        ("rel_cutoff", C.POINTER(C.c_double)),
        ("tax_files", C.POINTER(C.c_char_p)),
        ("n_levels", C.c_uint32),
        ("rel_filter", C.POINTER(C.c_double)),
        ("fpr_query", C.POINTER(C.c_double)),
        ("skip_lca", C.c_int),
        ("tax_root_node", C.c_char_p),
        ("output_lca", C.c_int),
        ("output_all", C.c_int),
        ("output_unclassified", C.c_int),
        ("output_single", C.c_int),
        ("device", C.c_int),
        ("host_threads", C.c_int),
        ("n_reads_chunk", C.c_int),
        ("quiet", C.c_int),
        ("cuda_stream", C.c_void_p),
        ("comm", C.c_void_p),
        ("sliced_ingest", C.c_int),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("consumed1", C.c_uint64),
        ("consumed2", C.c_uint64),
        ("parse_error", C.c_int),
        ("match_off", C.POINTER(C.c_uint64)),
        ("match_target", C.POINTER(C.c_uint32)),
        ("match_count", C.POINTER(C.c_uint32)),
        ("read_level", C.POINTER(C.c_uint8)),
        ("n_hashes", C.POINTER(C.c_uint32)),
        ("n_classified", C.c_uint64),
        ("n_levels", C.c_uint32),
        ("all_text", C.POINTER(C.c_void_p)),
        ("all_len", C.POINTER(C.c_uint64)),
        ("one_text", C.POINTER(C.c_void_p)),
        ("one_len", C.POINTER(C.c_uint64)),
        ("unc_text", C.c_void_p),
        ("unc_len", C.c_uint64),
        ("ms_h2d", C.c_float),
        ("ms_index", C.c_float),
        ("ms_minimiser", C.c_float),
        ("ms_count", C.c_float),
        ("ms_sort", C.c_float),
        ("ms_d2h", C.c_float),
        ("ms_host_index", C.c_double),
        ("ms_host_finish", C.c_double),
        ("ms_total", C.c_double),
        ("n_minimisers", C.c_uint64),
        ("count_kernel_bytes", C.c_uint64),
        ("n_kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("ms_finish_device", C.c_float),
        ("levels_on_device", C.c_uint32),
        ("ms_exchange", C.c_float),
        ("exchanged_bytes", C.c_uint64),
    ]


class OutputFds(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("all_fd", C.POINTER(C.c_int)), ("one_fd", C.POINTER(C.c_int)), ("unc_fd", C.c_int)]


class FilesResult(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_records", "n_classified", "n_blocks", "bytes_read1", "bytes_read2")] + [("parse_error", C.c_int), ("is_gzip", C.c_int)] + [
        (n, C.c_double) for n in ("ms_open", "ms_read_wait", "ms_submit", "ms_collect", "ms_write")
    ]


class BuildFileStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_sequences", "n_skipped", "n_bases", "n_hashes_total", "n_unique")] + [("parse_error", C.c_int)]


class ReassignResult(C.Structure):
    _fields_ = [
        ("n_groups", C.c_uint32),
        ("group_label", C.POINTER(C.c_char_p)),
        ("one_text", C.POINTER(C.c_void_p)),
        ("one_len", C.POINTER(C.c_uint64)),
        ("iterations", C.POINTER(C.c_uint32)),
        ("reassigned_reads", C.POINTER(C.c_uint64)),
        ("rep_text", C.c_void_p),
        ("rep_len", C.c_uint64),
    ]


class Totals(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("input_seqs", "seqs_processed", "seqs_skipped_big", "seqs_skipped_small", "length_processed", "kmers_processed", "seqs_classified", "kmers_matches", "kmers_from_classified_seqs", "matches", "seqs_unique", "discarded_matches_filter", "discarded_matches_fprquery")]


# every symbol include/ganon_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gnb_last_error": (C.c_char_p, []),
    "gnb_abi_version": (C.c_int, []),
    "gnb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gnb_db_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gnb_db_open_paged": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, C.POINTER(_P)]),
    "gnb_db_page_out": (C.c_int, [_P, C.c_uint64]),
    "gnb_db_info": (C.c_int, [_P, C.POINTER(DbInfo)]),
    "gnb_db_target": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gnb_db_free": (None, [_P]),
    "gnb_db_create": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_P)]),
    "gnb_db_create_sharded": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gnb_db_fill_random": (C.c_int, [_P, C.c_uint64, C.c_int]),
    "gnb_db_emplace": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "gnb_db_set_targets": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_char_p), _P, _P, C.c_uint64]),
    "gnb_db_set_fp": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "gnb_db_read_words": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "gnb_db_save": (C.c_int, [_P, C.c_char_p]),
    "gnb_db_create_hibf": (C.c_int, [C.c_uint64, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, C.c_uint64, C.POINTER(C.c_char_p), C.c_double, C.c_int, C.POINTER(_P)]),
    "gnb_db_emplace_ibf": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_uint64]),
    "gnb_build_file_hashes": (C.c_int, [C.c_int, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_P), C.POINTER(BuildFileStats)]),
    "gnb_hash_set_data": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gnb_hash_set_free": (None, [_P]),
    "gnb_comm_unique_id": (C.c_int, [_P, C.c_uint64]),
    "gnb_comm_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gnb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gnb_comm_free": (None, [_P]),
    "gnb_minimisers": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "gnb_minimisers_batch": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, _P, _P, C.c_uint64, _P, _P, C.c_uint64]),
    "gnb_db_bulk_count": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_uint64, _P]),
    "gnb_session_create": (C.c_int, [C.POINTER(SessionConfig), C.POINTER(_P)]),
    "gnb_session_free": (None, [_P]),
    "gnb_session_classify": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.POINTER(BatchResult)]),
    "gnb_session_stage": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "gnb_session_run_staged": (C.c_int, [_P, C.POINTER(BatchResult)]),
    "gnb_session_finish_staged": (C.c_int, [_P, C.c_uint32, C.POINTER(BatchResult)]),
    "gnb_session_submit": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.POINTER(BatchResult)]),
    "gnb_session_collect": (C.c_int, [_P, C.POINTER(BatchResult)]),
    "gnb_session_in_flight": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "gnb_session_run_level": (C.c_int, [_P, C.c_uint32]),
    "gnb_session_level_tuples": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gnb_session_set_level_tuples": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
    "gnb_session_finish_level": (C.c_int, [_P, C.c_uint32]),
    "gnb_session_collect_staged": (C.c_int, [_P, C.c_uint32, C.POINTER(BatchResult)]),
    "gnb_session_run_level_device": (C.c_int, [_P, C.c_uint32]),
    "gnb_session_level_tuples_device": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gnb_session_set_level_tuples_device": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64]),
    "gnb_session_finish_level_device": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "gnb_session_hibf_rounds": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.POINTER(C.c_uint32)]),
    "gnb_session_staged_timings": (C.c_int, [_P, C.POINTER(BatchResult)]),
    "gnb_reads_file_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "gnb_reads_file_read": (C.c_int64, [_P, _P, C.c_uint64]),
    "gnb_reads_file_is_gzip": (C.c_int, [_P]),
    "gnb_reads_file_close": (None, [_P]),
    "gnb_session_classify_files": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_char_p, C.POINTER(OutputFds), C.c_uint64, C.c_int, C.POINTER(FilesResult)]),
    "gnb_host_register": (C.c_int, [_P, C.c_uint64]),
    "gnb_host_unregister": (C.c_int, [_P]),
    "gnb_session_level_count": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "gnb_session_level_label": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_char_p)]),
    "gnb_session_node_name": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p)]),
    "gnb_session_report": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gnb_session_stats": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gnb_session_totals": (C.c_int, [_P, C.c_uint32, C.c_int, C.POINTER(Totals)]),
    "gnb_session_keep_matches": (C.c_int, [_P, C.c_int]),
    "gnb_session_reassign": (C.c_int, [_P, C.c_uint32, C.c_double, C.c_uint32, C.POINTER(ReassignResult)]),
}

_LIB = None


def build(force: bool = False) -> str:
    """Compile the shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-s", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    subprocess.check_call(args)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libganon_b200.so is not built (run `make -C ganon_b200/csrc` or __graft_entry__.build()); " "there is no fallback implementation")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != GNB_OK:
        raise GnbError(rc, lib().gnb_last_error().decode(errors="replace"))
