"""ganon database file formats (.ibf flat IBF, .hibf raptor HIBF) -- pure numpy reader / writer.

The product's loader is the C++ parser inside libganon_b200.so (csrc/db_file.cpp); this module is
the host-side mirror used for inspection, for writing synthetic databases (bench / tests) and for
cross-checking the native parser.  Layouts (little-endian cereal binary, no magic, no padding):

.ibf  -- writer ``save_filter`` src/ganon-build/GanonBuild.cpp:251-288, reader GanonClassify.cpp:949-986,
         ``IBFConfig`` src/utils/include/utils/IBFConfig.hpp:18-40, IBF ``serialize``
         seqan3 interleaved_bloom_filter.hpp:561-571, ``sdsl::bit_vector`` int_vector.hpp:2029-2063.
.hibf -- raptor 3.0.1 index, reader GanonClassify.cpp:875-938, HIBF ``serialize``
         hierarchical_interleaved_bloom_filter.hpp:163-169,293-298.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import BinaryIO, List, Optional, Sequence, Tuple

import numpy as np

GANON_VERSION = (2, 4, 1)


def _countl_zero64(x: int) -> int:
    return 64 - int(x).bit_length()


@dataclass
class IBF:
    """seqan3::interleaved_bloom_filter<uncompressed>: row-major [row][bin_word] 64-bit words, LSB-first."""

    bins: int
    bin_size: int
    hash_funs: int
    data: Optional[np.ndarray] = None  # uint64[bin_size * bin_words]; None = header only
    technical_bins: int = 0
    hash_shift: int = 0
    bin_words: int = 0

    def __post_init__(self):
        if not self.bin_words:
            self.bin_words = (self.bins + 63) >> 6
        if not self.technical_bins:
            self.technical_bins = self.bin_words << 6
        if not self.hash_shift:
            self.hash_shift = _countl_zero64(self.bin_size)

    @property
    def n_words(self) -> int:
        return self.bin_size * self.bin_words


@dataclass
class IBFFile:
    """Everything a ganon .ibf holds."""

    ibf: IBF
    kmer_size: int
    window_size: int
    max_hashes_bin: int
    hashes_count: List[Tuple[str, int]]  # (target, number of minimiser hashes)
    bin_map: List[Tuple[int, str]]  # (technical bin, target), unsorted in the reference
    max_fp: float = 0.05
    true_max_fp: float = 0.0
    true_avg_fp: float = 0.0
    version: Tuple[int, int, int] = GANON_VERSION
    data_offset: int = 0  # byte offset of the bitvector payload inside the file


def _w_str(f: BinaryIO, s: str) -> None:
    b = s.encode()
    f.write(struct.pack("<Q", len(b)))
    f.write(b)


def _r_str(f: BinaryIO) -> str:
    (n,) = struct.unpack("<Q", f.read(8))
    return f.read(n).decode()


def _write_ibf_body(f: BinaryIO, ibf: IBF, data_chunks=None) -> None:
    f.write(struct.pack("<6Q", ibf.bins, ibf.technical_bins, ibf.bin_size, ibf.hash_shift, ibf.bin_words, ibf.hash_funs))
    n_bits = ibf.technical_bins * ibf.bin_size
    f.write(struct.pack("<BfQ", 1, 1.5, n_bits))
    if data_chunks is not None:
        total = 0
        for c in data_chunks:
            c = np.ascontiguousarray(c, dtype="<u8")
            f.write(memoryview(c).cast("B"))
            total += c.size
        assert total == ibf.n_words, (total, ibf.n_words)
    else:
        assert ibf.data is not None and ibf.data.size == ibf.n_words
        f.write(memoryview(np.ascontiguousarray(ibf.data, dtype="<u8")).cast("B"))


def _read_ibf_body(f: BinaryIO, load_data: bool = True) -> IBF:
    bins, tech, bin_size, shift, words, funs = struct.unpack("<6Q", f.read(48))
    width, _growth, n_bits = struct.unpack("<BfQ", f.read(13))
    assert width == 1 and n_bits == tech * bin_size, (width, n_bits, tech, bin_size)
    n_words = (n_bits + 63) >> 6
    data = None
    if load_data:
        data = np.frombuffer(f.read(n_words * 8), dtype="<u8").copy()
        assert data.size == n_words
    else:
        f.seek(n_words * 8, 1)
    return IBF(bins=bins, bin_size=bin_size, hash_funs=funs, data=data, technical_bins=tech, hash_shift=shift, bin_words=words)


def write_ibf(path: str, db: IBFFile, data_chunks=None) -> None:
    """Write a flat ganon .ibf.  ``data_chunks``: optional iterable of uint64 arrays streamed as the payload."""
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", *db.version))
        f.write(
            struct.pack(
                "<QQBBHQddd",
                db.ibf.bins,
                db.max_hashes_bin,
                db.ibf.hash_funs,
                db.kmer_size,
                db.window_size,
                db.ibf.bin_size,
                db.max_fp,
                db.true_max_fp,
                db.true_avg_fp,
            )
        )
        f.write(struct.pack("<Q", len(db.hashes_count)))
        for t, c in db.hashes_count:
            _w_str(f, t)
            f.write(struct.pack("<Q", c))
        f.write(struct.pack("<Q", len(db.bin_map)))
        for b, t in db.bin_map:
            f.write(struct.pack("<Q", b))
            _w_str(f, t)
        _write_ibf_body(f, db.ibf, data_chunks)


def read_ibf(path: str, load_data: bool = True) -> IBFFile:
    with open(path, "rb") as f:
        version = struct.unpack("<3i", f.read(12))
        n_bins, max_hashes_bin, hf, k, w, bin_size_bits, max_fp, tmax, tavg = struct.unpack("<QQBBHQddd", f.read(52))
        (n,) = struct.unpack("<Q", f.read(8))
        hc = []
        for _ in range(n):
            t = _r_str(f)
            (c,) = struct.unpack("<Q", f.read(8))
            hc.append((t, c))
        (m,) = struct.unpack("<Q", f.read(8))
        bm = []
        for _ in range(m):
            (b,) = struct.unpack("<Q", f.read(8))
            bm.append((b, _r_str(f)))
        off = f.tell() + 48 + 13
        ibf = _read_ibf_body(f, load_data)
        assert ibf.bin_size == bin_size_bits and ibf.hash_funs == hf
    return IBFFile(ibf, k, w, max_hashes_bin, hc, bm, max_fp, tmax, tavg, tuple(version), off)


# ----------------------------------------------------------------------------- HIBF


@dataclass
class HIBFFile:
    """raptor 3.0.1 HIBF index as read by ganon-classify --hibf."""

    window_size: int
    kmer_size: int
    ibfs: List[IBF]
    next_ibf_id: List[List[int]]  # [ibf][technical bin] -> child ibf (or own index)
    user_bin_filenames: List[str]
    bin_to_user: List[List[int]]  # ibf_bin_to_filename_position; <0 = merged bin
    bin_path: List[List[str]]  # user bin -> list of file paths (target name = basename of first)
    fpr: float = 0.05
    version: int = 1
    parts: int = 1
    compressed: bool = False
    is_hibf: bool = True
    data_offsets: List[int] = field(default_factory=list)


def _w_vec_i64(f: BinaryIO, v: Sequence[int]) -> None:
    f.write(struct.pack("<Q", len(v)))
    f.write(np.asarray(v, dtype="<i8").tobytes())


def _r_vec_i64(f: BinaryIO) -> List[int]:
    (n,) = struct.unpack("<Q", f.read(8))
    return np.frombuffer(f.read(8 * n), dtype="<i8").tolist()


def write_hibf(path: str, db: HIBFFile) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<IQ", db.version, db.window_size))
        # seqan3::shape = dynamic_bitset<58>: serialised as one u64 holding {size:6 bits | bits:58} -- see
        # dynamic_bitset.hpp:1963-1972 (archive(data.size), archive(data.bits))
        f.write(struct.pack("<QQ", db.kmer_size, (1 << db.kmer_size) - 1))
        f.write(struct.pack("<BB", db.parts, 1 if db.compressed else 0))
        f.write(struct.pack("<Q", len(db.bin_path)))
        for paths in db.bin_path:
            f.write(struct.pack("<Q", len(paths)))
            for p in paths:
                _w_str(f, p)
        f.write(struct.pack("<dB", db.fpr, 1 if db.is_hibf else 0))
        f.write(struct.pack("<Q", len(db.ibfs)))
        for ibf in db.ibfs:
            _write_ibf_body(f, ibf)
        f.write(struct.pack("<Q", len(db.next_ibf_id)))
        for v in db.next_ibf_id:
            _w_vec_i64(f, v)
        f.write(struct.pack("<Q", len(db.user_bin_filenames)))
        for s in db.user_bin_filenames:
            _w_str(f, s)
        f.write(struct.pack("<Q", len(db.bin_to_user)))
        for v in db.bin_to_user:
            _w_vec_i64(f, v)


def read_hibf(path: str, load_data: bool = True) -> HIBFFile:
    with open(path, "rb") as f:
        version, window = struct.unpack("<IQ", f.read(12))
        shape_size, shape_bits = struct.unpack("<QQ", f.read(16))
        parts, compressed = struct.unpack("<BB", f.read(2))
        (n,) = struct.unpack("<Q", f.read(8))
        bin_path = []
        for _ in range(n):
            (m,) = struct.unpack("<Q", f.read(8))
            bin_path.append([_r_str(f) for _ in range(m)])
        fpr, is_hibf = struct.unpack("<dB", f.read(9))
        (n_ibf,) = struct.unpack("<Q", f.read(8))
        ibfs, offs = [], []
        for _ in range(n_ibf):
            offs.append(f.tell() + 48 + 13)
            ibfs.append(_read_ibf_body(f, load_data))
        (n,) = struct.unpack("<Q", f.read(8))
        nxt = [_r_vec_i64(f) for _ in range(n)]
        (n,) = struct.unpack("<Q", f.read(8))
        names = [_r_str(f) for _ in range(n)]
        (n,) = struct.unpack("<Q", f.read(8))
        pos = [_r_vec_i64(f) for _ in range(n)]
        assert f.read(1) == b"", "trailing bytes in .hibf"
    return HIBFFile(window, bin(shape_bits).count("1"), ibfs, nxt, names, pos, bin_path, fpr, version, parts, bool(compressed), bool(is_hibf), offs)


def hibf_target_name(path: str) -> str:
    """Target name of a user bin, GanonClassify.cpp:916-928."""
    f = path.rsplit("/", 1)[-1]
    i = f.find(".minimiser")
    if i >= 0:
        f = f[:i]
    return f.replace("|||", ".").replace("---", " ")
