"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/liboracle.so (ganon_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (ganon_b200/) never does.  Parity status: pinned, see ganon_oracle.h.

The per-read control flow below (classify_level) restates GanonClassify.cpp:630-832 ``classify()``.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class GoIbf(C.Structure):
    _fields_ = [
        ("bins", C.c_uint64),
        ("technical_bins", C.c_uint64),
        ("bin_size", C.c_uint64),
        ("hash_shift", C.c_uint64),
        ("bin_words", C.c_uint64),
        ("hash_funs", C.c_uint64),
        ("data", C.c_void_p),
    ]


class GoHibf(C.Structure):
    _fields_ = [
        ("n_ibf", C.c_size_t),
        ("ibfs", C.POINTER(GoIbf)),
        ("next_ibf_id", C.POINTER(C.c_void_p)),
        ("bin_to_user", C.POINTER(C.c_void_p)),
        ("n_user_bins", C.c_size_t),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ganon_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.go_adjust_seed.restype = C.c_uint64
        L.go_adjust_seed.argtypes = [C.c_uint]
        L.go_dna4_rank.restype = C.c_uint
        L.go_dna4_rank.argtypes = [C.c_ubyte]
        L.go_dna15_valid.restype = C.c_int
        L.go_dna15_valid.argtypes = [C.c_ubyte]
        L.go_minimiser_hash.restype = C.c_size_t
        L.go_minimiser_hash.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_uint64, C.c_void_p]
        L.go_ibf_row.restype = C.c_uint64
        L.go_ibf_row.argtypes = [C.POINTER(GoIbf), C.c_uint64, C.c_uint]
        L.go_ibf_bulk_count.restype = None
        L.go_ibf_bulk_count.argtypes = [C.POINTER(GoIbf), C.c_void_p, C.c_size_t, C.c_void_p]
        L.go_ibf_emplace.restype = None
        L.go_ibf_emplace.argtypes = [C.POINTER(GoIbf), C.c_void_p, C.c_uint64, C.c_uint64]
        L.go_threshold_cutoff.restype = C.c_uint64
        L.go_threshold_cutoff.argtypes = [C.c_uint64, C.c_double]
        L.go_threshold_filter.restype = C.c_uint64
        L.go_threshold_filter.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        L.go_select_matches_ibf.restype = None
        L.go_select_matches_ibf.argtypes = [C.POINTER(GoIbf)] + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint64] + [C.c_void_p] * 5
        L.go_hibf_bulk_count.restype = None
        L.go_hibf_bulk_count.argtypes = [C.POINTER(GoHibf), C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]
        L.go_fpr_query_q.restype = C.c_double
        L.go_fpr_query_q.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        L.go_target_fpr.restype = C.c_double
        L.go_target_fpr.argtypes = [C.c_uint64, C.c_uint, C.c_uint64, C.c_uint64]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- primitives


def adjust_seed(k: int) -> int:
    return lib().go_adjust_seed(k)


def minimiser_hash(seq: bytes, k: int, w: int, seed: Optional[int] = None) -> np.ndarray:
    """seqan3::views::minimiser_hash(ungapped{k}, window_size{w}, seed); default seed = adjust_seed(k)."""
    if isinstance(seq, str):
        seq = seq.encode()
    if seed is None:
        seed = adjust_seed(k)
    out = np.empty(max(len(seq) - k + 1, 1), dtype=np.uint64)
    n = lib().go_minimiser_hash(seq, len(seq), k, w, seed, _p(out))
    return out[:n].copy()


def read_hashes(seq1: bytes, seq2: Optional[bytes], k: int, w: int) -> Optional[np.ndarray]:
    """Hash list of a read (pair) as built in GanonClassify.cpp:690-700; None = skipped (shorter than window)."""
    if len(seq1) < w:
        return None
    h = minimiser_hash(seq1, k, w)
    if seq2 is not None and len(seq2) >= w:
        h = np.concatenate([h, minimiser_hash(seq2, k, w)])
    return h


class OracleIBF:
    """Host-resident flat IBF + its target map, the oracle's view of a loaded .ibf."""

    def __init__(self, bins: int, bin_size: int, hash_funs: int, data: Optional[np.ndarray] = None):
        self.bin_words = (bins + 63) >> 6
        self.technical_bins = self.bin_words * 64
        self.bins, self.bin_size, self.hash_funs = bins, bin_size, hash_funs
        self.hash_shift = 64 - int(bin_size).bit_length()
        self.data = np.zeros(self.bin_words * bin_size, dtype=np.uint64) if data is None else np.ascontiguousarray(data, dtype=np.uint64)
        assert self.data.size == self.bin_words * bin_size
        self.c = GoIbf(bins, self.technical_bins, bin_size, self.hash_shift, self.bin_words, hash_funs, self.data.ctypes.data)

    def rows(self, value: int) -> List[int]:
        return [lib().go_ibf_row(C.byref(self.c), int(value), i) for i in range(self.hash_funs)]

    def emplace(self, value: int, bin_: int) -> None:
        lib().go_ibf_emplace(C.byref(self.c), _p(self.data), int(value), int(bin_))

    def emplace_many(self, values: Iterable[int], bin_: int) -> None:
        L, c, d = lib(), C.byref(self.c), _p(self.data)
        for v in values:
            L.go_ibf_emplace(c, d, int(v), int(bin_))

    def bulk_count(self, hashes: np.ndarray) -> np.ndarray:
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        counts = np.empty(self.technical_bins, dtype=np.uint16)
        lib().go_ibf_bulk_count(C.byref(self.c), _p(hashes), hashes.size, _p(counts))
        return counts


class OracleHIBF:
    def __init__(self, ibfs: Sequence[OracleIBF], next_ibf_id: Sequence[Sequence[int]], bin_to_user: Sequence[Sequence[int]], n_user_bins: int):
        self.ibfs = list(ibfs)
        self._c_ibfs = (GoIbf * len(ibfs))(*[i.c for i in ibfs])
        self._nxt = [np.ascontiguousarray(v, dtype=np.int64) for v in next_ibf_id]
        self._pos = [np.ascontiguousarray(v, dtype=np.int64) for v in bin_to_user]
        self._nxt_p = (C.c_void_p * len(ibfs))(*[v.ctypes.data for v in self._nxt])
        self._pos_p = (C.c_void_p * len(ibfs))(*[v.ctypes.data for v in self._pos])
        self.n_user_bins = n_user_bins
        self.c = GoHibf(len(ibfs), self._c_ibfs, self._nxt_p, self._pos_p, n_user_bins)

    def bulk_count(self, hashes: np.ndarray, threshold: int) -> np.ndarray:
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        res = np.empty(self.n_user_bins, dtype=np.uint16)
        lib().go_hibf_bulk_count(C.byref(self.c), _p(hashes), hashes.size, int(threshold), _p(res))
        return res


# ----------------------------------------------------------------------------- read files


def parse_reads(path: str) -> List[Tuple[bytes, bytes]]:
    """Minimal FASTA/FASTQ (optionally gz) reader: (id = whole header line, sequence).

    Mirrors what ganon-classify gets from seqan3 (format_fastq.hpp:105-267 / format_fasta.hpp,
    truncate_ids=false).  Raises ValueError on a character that is not legal for dna15.
    """
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    out: List[Tuple[bytes, bytes]] = []
    i, n = 0, len(lines)
    L = lib()
    while i < n:
        ln = lines[i].rstrip(b"\r")
        if not ln:
            i += 1
            continue
        if ln[:1] == b">":
            rid = ln[1:]
            i += 1
            seq = []
            while i < n and lines[i][:1] != b">":
                seq.append(lines[i].strip())
                i += 1
            s = b"".join(seq)
        elif ln[:1] == b"@":
            rid = ln[1:]
            i += 1
            seq = []
            while i < n and lines[i][:1] != b"+":
                seq.append(lines[i].strip())
                i += 1
            s = b"".join(seq)
            i += 1  # '+' line
            q = 0
            while i < n and q < len(s):
                q += len(lines[i].strip())
                i += 1
        else:
            raise ValueError("unrecognised record start: %r" % ln[:20])
        for ch in s:
            if not L.go_dna15_valid(ch):
                raise ValueError("illegal character %r" % chr(ch))
        out.append((rid, s))
    return out


# ----------------------------------------------------------------------------- classification of one hierarchy level


class OracleFilter:
    """One database of a hierarchy level (flat IBF or HIBF) with its target table."""

    def __init__(self, ibf, targets: Sequence[str], target_bins: Sequence[Sequence[int]], target_fpr: Sequence[float], rel_cutoff: float, k: int, w: int):
        self.ibf = ibf
        self.is_hibf = isinstance(ibf, OracleHIBF)
        self.targets = list(targets)
        self.target_bins = [list(b) for b in target_bins]
        self.target_fpr = np.asarray(target_fpr, dtype=np.float64)
        self.rel_cutoff = rel_cutoff
        self.k, self.w = k, w
        off = np.zeros(len(targets) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(b) for b in target_bins])
        self.off = off
        self.flat_bins = np.asarray([x for b in target_bins for x in b], dtype=np.uint64)

    @staticmethod
    def from_ibf_file(dbf, rel_cutoff: float) -> "OracleFilter":
        """Build from ganon_b200.formats.IBFFile, as load_filter/load_files do (GanonClassify.cpp:949-986,1007-1039)."""
        ibf = OracleIBF(dbf.ibf.bins, dbf.ibf.bin_size, dbf.ibf.hash_funs, dbf.ibf.data)
        tmap: Dict[str, List[int]] = {}
        for b, t in dbf.bin_map:
            tmap.setdefault(t, []).append(b)
        counts = dict(dbf.hashes_count)
        targets = list(tmap)
        L = lib()
        # target_fpr[target] defaults to 0.0 when the target is absent from hashes_count (operator[] on the map)
        fpr = [L.go_target_fpr(dbf.ibf.bin_size, dbf.ibf.hash_funs, dbf.max_hashes_bin, counts[t]) if t in counts else 0.0 for t in targets]
        return OracleFilter(ibf, targets, [tmap[t] for t in targets], fpr, rel_cutoff, dbf.kmer_size, dbf.window_size)


def classify_level(filters: Sequence[OracleFilter], reads: Sequence[Tuple[bytes, bytes, Optional[bytes]]], rel_filter: float, fpr_query: float):
    """``classify()`` GanonClassify.cpp:630-832 for one hierarchy level, without LCA / report bookkeeping.

    reads: (id, seq1, seq2-or-None).  Returns a list, one item per read:
      dict(n_hashes, skipped, max, min, matches=[(target, count)...] after cutoff+rel_filter+fpr_query,
           discarded_filter=[targets], discarded_fpr=[targets])
    """
    L = lib()
    k, w = filters[0].k, filters[0].w
    # level-wide target ids (TMatches is keyed by target name across filters)
    gid: Dict[str, int] = {}
    for f in filters:
        for t in f.targets:
            gid.setdefault(t, len(gid))
    names = list(gid)
    fgid = [np.asarray([gid[t] for t in f.targets], dtype=np.uint32) for f in filters]
    best = np.zeros(len(names), dtype=np.uint64)
    bfpr = np.zeros(len(names), dtype=np.float64)
    out = []
    for rid, s1, s2 in reads:
        res = dict(id=rid, n_hashes=0, skipped=None, max=0, min=0, matches=[], discarded_filter=[], discarded_fpr=[])
        h = read_hashes(s1, s2, k, w)
        if h is None:
            res["skipped"] = "small"
            out.append(res)
            continue
        n = int(h.size)
        res["n_hashes"] = n
        if n > 65535:
            res["skipped"] = "big"
            out.append(res)
            continue
        best[:] = 0
        mx, mn = C.c_uint64(0), C.c_uint64(n)
        for f, fg in zip(filters, fgid):
            cutoff = L.go_threshold_cutoff(n, f.rel_cutoff)
            if not f.is_hibf:
                counts = np.empty(f.ibf.technical_bins, dtype=np.uint16)
                L.go_select_matches_ibf(C.byref(f.ibf.c), _p(f.off), _p(f.flat_bins), _p(fg), _p(f.target_fpr), len(f.targets), _p(h), n, cutoff, _p(best), _p(bfpr), C.byref(mx), C.byref(mn), _p(counts))
            else:
                # select_matches(THIBF) GanonClassify.cpp:543-577
                cnt = f.ibf.bulk_count(h, cutoff)
                for t in range(len(f.targets)):
                    c = int(cnt[f.target_bins[t][0]])
                    if c > 0:
                        c = min(c, n)
                        g = fg[t]
                        if c > best[g]:
                            best[g] = c
                            bfpr[g] = f.target_fpr[t]
                            mx.value = max(mx.value, c)
                            mn.value = min(mn.value, c)
        res["max"], res["min"] = mx.value, mn.value
        if mx.value > 0:
            thr = L.go_threshold_filter(mx.value, mn.value, rel_filter)
            for g in np.nonzero(best)[0]:
                c = int(best[g])
                if c >= thr:
                    if fpr_query < 1.0 and L.go_fpr_query_q(n, c, float(bfpr[g])) > fpr_query:
                        res["discarded_fpr"].append(names[g])
                        continue
                    res["matches"].append((names[g], c))
                else:
                    res["discarded_filter"].append(names[g])
        out.append(res)
    return out


def all_lines(results) -> List[str]:
    """The `.all` file content (unordered in the reference -> compare sorted)."""
    return sorted("%s\t%s\t%d" % (r["id"].decode(), t, c) for r in results for t, c in r["matches"])


# ---- taxonomy + LCA (A7): utils/LCA.hpp:38-174, GanonClassify.cpp:615-627, 988-1005, 1324-1371 ---------------------------
def load_tax(path: str) -> Dict[str, str]:
    """``load_tax`` GanonClassify.cpp:988-1005: node <tab> parent <tab> rank <tab> name; later lines replace earlier ones.
    Only the parent matters for the LCA."""
    tax: Dict[str, str] = {}
    with open(path) as f:
        for line in f:
            fields = line.rstrip("\n").split("\t")
            tax[fields[0]] = fields[1]
    return tax


def merge_tax(taxes: Sequence[Dict[str, str]]) -> Dict[str, str]:
    """``merge_tax`` GanonClassify.cpp:1324-1342: the first filter's entry of a node wins (map::insert)."""
    merged = dict(taxes[0])
    for t in taxes[1:]:
        for node, parent in t.items():
            merged.setdefault(node, parent)
    return merged


def validate_targets_tax(tax: Dict[str, str], filters: Sequence["OracleFilter"], root: str = "1") -> Dict[str, str]:
    """``validate_targets_tax`` GanonClassify.cpp:1344-1362: targets without a tax entry hang under the root node."""
    for f in filters:
        for t in f.targets:
            tax.setdefault(t, root)
    return tax


class OracleLCA:
    """``pre_process_lca`` (GanonClassify.cpp:1364-1371: one edge parent -> node per tax entry, Euler walk from the root)
    and ``LCA::getLCA`` (utils/LCA.hpp:150-174: pairwise fold over the targets).  The reference answers the pairwise
    query with an Euler tour + sparse-table RMQ (LCA.hpp:60-148); the lowest common ancestor of two nodes of a tree is
    unique, so this restatement walks parent pointers with the depths of the same depth-first search instead."""

    def __init__(self, tax: Dict[str, str], root: str = "1"):
        children: Dict[str, List[str]] = {}
        for node, parent in tax.items():
            children.setdefault(parent, []).append(node)
        self.parent: Dict[str, str] = {}
        self.depth: Dict[str, int] = {root: 0}
        stack = [root]
        while stack:  # depthFirstSearch LCA.hpp:60-82 (only nodes reachable from the root get a first appearance)
            cur = stack.pop()
            for ch in children.get(cur, []):
                if ch not in self.depth:
                    self.depth[ch] = self.depth[cur] + 1
                    self.parent[ch] = cur
                    stack.append(ch)

    def pair(self, u: str, v: str) -> str:
        du, dv = self.depth[u], self.depth[v]
        while du > dv:
            u, du = self.parent[u], du - 1
        while dv > du:
            v, dv = self.parent[v], dv - 1
        while u != v:
            u, v = self.parent[u], self.parent[v]
        return u

    def get_lca(self, targets: Sequence[str]) -> str:
        assert len(targets) > 1  # LCA.hpp:168
        lca = self.pair(targets[0], targets[1])
        for t in targets[2:]:
            lca = self.pair(lca, t)
        return lca


def one_lines(results, lca: Optional[OracleLCA]) -> List[str]:
    """The `.one` file (--output-lca) of a level: the single match of a read, or the LCA of its matches with the read's
    maximum count (GanonClassify.cpp:770-786, 615-627).  Compare sorted."""
    out = []
    for r in results:
        m = r["matches"]
        if len(m) == 1:
            out.append("%s\t%s\t%d" % (r["id"].decode(), m[0][0], m[0][1]))
        elif len(m) > 1:
            out.append("%s\t%s\t%d" % (r["id"].decode(), lca.get_lca([t for t, _ in m]), r["max"]))
    return sorted(out)
