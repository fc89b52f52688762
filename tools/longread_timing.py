#!/usr/bin/env python
"""Time K2 on batches of long reads: the default (K2t over segments of 512 windows where the batch holds long reads) against
GANON_B200_K2=thread / =warp (one thread / one warp walks a whole read).  Not a bench line: a diagnosis of minimisers_segmented()
(kernels.cu)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(read_len, n_reads, one_long=0):
    import numpy as np

    from ganon_b200.classify import Database, Session

    rng = np.random.default_rng(5)
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n_reads, read_len))
    recs = []
    qual = b"I" * read_len
    for i in range(n_reads):
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, seq[i].tobytes(), qual))
    if one_long:  # one very long read in the middle of the batch
        big = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=one_long).tobytes()
        recs.insert(n_reads // 2, b"@big\n%s\n+\n%s\n" % (big, b"I" * one_long))
    text = b"".join(recs)
    db = Database.create(256, 1 << 20, 4, 19, 31)
    db.fill_random(3, 2)
    s = Session([db], [0.0], [1.0], [1.0])
    best = None
    for _ in range(4):
        r = s.classify(text, final=True)
        t = (r.ms_minimiser, r.ms_count, r.ms_total)
        best = t if best is None or t[0] < best[0] else best
    print(json.dumps({"k2": os.environ.get("GANON_B200_K2", "default"), "read_len": read_len, "n_reads": n_reads, "one_long": one_long, "ms_hash": best[0], "ms_count": best[1], "ms_total": best[2],
                      "gbases_per_s_k2": (read_len * n_reads + one_long) / best[0] / 1e6}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    else:
        for read_len, n_reads, one_long in ((10_000, 3000, 0), (50_000, 600, 0), (1_000, 30_000, 0), (3_000, 10_000, 0), (150, 200_000, 0), (150, 200_000, 1_000_000)):
            for k2 in ("", "thread", "warp"):
                env = dict(os.environ)
                if k2:
                    env["GANON_B200_K2"] = k2
                subprocess.run([sys.executable, os.path.abspath(__file__), str(read_len), str(n_reads), str(one_long)], env=env, check=False)
