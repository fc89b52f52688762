#!/usr/bin/env python
"""Throughput of the library's gzip reader alone (gnb_reads_file_*: no GPU work) on a synthetic single-member FASTQ .gz,
by worker count.  A diagnosis for the `cli.gz` figure of the bench line, not a bench value."""
import ctypes as C
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import bench  # noqa: E402
from ganon_b200 import _lib  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n_reads, 150))
    rec = np.empty((n_reads, 8 + 8 + 1 + 150 + 3 + 150 + 1), dtype=np.uint8)  # "@r" + 14 digits + "\n" seq "\n+\n" qual "\n"
    ids = np.char.zfill(np.arange(n_reads).astype("U14"), 14).astype("S14").view(np.uint8).reshape(n_reads, 14)
    rec[:, 0:2] = np.frombuffer(b"@r", dtype=np.uint8)
    rec[:, 2:16] = ids
    rec[:, 16] = 10
    rec[:, 17:167] = seq
    rec[:, 167:170] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 170:320] = ord("I")
    rec[:, 320] = 10
    with tempfile.TemporaryDirectory(dir=os.environ.get("TMPDIR", "/tmp")) as d:
        path = os.path.join(d, "reads.fq.gz")
        t0 = time.perf_counter()
        bench.write_single_member_gzip(path, [rec])
        print(json.dumps({"fastq_bytes": int(rec.size), "gz_bytes": os.path.getsize(path), "compress_s": round(time.perf_counter() - t0, 2), "host_cpus": os.cpu_count()}), flush=True)
        L = _lib.lib()
        buf = C.create_string_buffer(64 << 20)
        for threads in (1, 2, 4, 8, 16, 24, 32):
            if threads > 2 * (os.cpu_count() or 1):
                break
            best = None
            for _ in range(2):
                h = C.c_void_p()
                t0 = time.perf_counter()
                assert L.gnb_reads_file_open(path.encode(), threads, C.byref(h)) == 0
                total = 0
                while True:
                    n = L.gnb_reads_file_read(h, buf, len(buf))
                    assert n >= 0, L.gnb_last_error()
                    if n == 0:
                        break
                    total += n
                t = time.perf_counter() - t0
                L.gnb_reads_file_close(h)
                assert total == rec.size
                best = t if best is None or t < best else best
            print(json.dumps({"threads": threads, "s": round(best, 3), "GB_per_s_out": round(rec.size / best / 1e9, 2), "M_reads_per_s": round(n_reads / best / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
