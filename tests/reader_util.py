"""The library's host record reader (ganon_b200/csrc/reads.cpp: FASTA, wrapped FASTQ, blanks, parse errors -- everything
the device indexer K1 hands back) compiled for the CPU (tests/native/reads_host.cpp) and driven the way the session drives
it: blocks of arbitrary size, the unconsumed tail carried into the next block, a parse error ends the file and loses the
chunk of --n-reads records being assembled (GanonClassify.cpp:1220-1287).

`differential_case` writes a randomly *formatted* reads file (same records, different dress: wrapped lines, blanks,
digits, CRLF, missing final newline, illegal letters ...), lets the UNMODIFIED reference binary classify it, and compares
with oracle classification of the records this reader returns.  Used by tests/test_reader_cpu.py."""
import ctypes as C
import os
import random
import shutil
import subprocess

from oracle import oracle as O
from tests import fuzz_util as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib(build_dir):
    global _LIB
    if _LIB is None:
        cxx = shutil.which("g++")
        if cxx is None:
            return None
        so = os.path.join(build_dir, "reads_host.so")
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", os.path.join(ROOT, "ganon_b200", "csrc"), "-o", so,
                               os.path.join(ROOT, "tests", "native", "reads_host.cpp")])
        L = C.CDLL(so)
        L.rh_index.restype = C.c_void_p
        L.rh_index.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint64]
        for f in ("rh_size", "rh_consumed", "rh_error_record"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.rh_error.restype = C.c_int
        L.rh_error.argtypes = [C.c_void_p]
        L.rh_error_msg.restype = C.c_char_p
        L.rh_error_msg.argtypes = [C.c_void_p]
        L.rh_get.restype = None
        L.rh_get.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        L.rh_free.restype = None
        L.rh_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def index_block(L, block: bytes, final: bool, max_records: int = 1 << 30):
    """-> (records [(id, seq)], consumed bytes, error record or None, message)."""
    h = L.rh_index(block, len(block), int(final), max_records)
    try:
        recs = []
        idp, seqp, idl, seql = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32()
        for i in range(L.rh_size(h)):
            L.rh_get(h, block, len(block), i, C.byref(idp), C.byref(idl), C.byref(seqp), C.byref(seql))
            recs.append((C.string_at(idp.value, idl.value), C.string_at(seqp.value, seql.value)))
        err = L.rh_error_record(h) if L.rh_error(h) else None
        return recs, L.rh_consumed(h), err, L.rh_error_msg(h).decode()
    finally:
        L.rh_free(h)


def read_file(L, data: bytes, block_bytes: int, n_reads: int):
    """The whole file through blocks of `block_bytes` as the session does (csrc/session.cpp, stage_block): records kept,
    whether a parse error ended the file."""
    out, pos, carry = [], 0, b""
    while True:
        chunk = data[pos : pos + block_bytes]
        pos += len(chunk)
        block = carry + chunk
        final = pos >= len(data)
        if not block:
            break
        recs, consumed, err, _msg = index_block(L, block, final)
        if err is not None and err <= len(recs):
            # record e fails while the reader looks one record ahead, i.e. inside the chunk that holds record e - 1: that chunk
            # is lost with everything after it (seqan3::views::chunk / std::views::take advance the file before they end)
            e = len(out) + err
            keep_abs = (e - 1) // n_reads * n_reads if e > 0 else 0
            out += recs[: max(0, keep_abs - len(out))]
            del out[keep_abs:]
            return out, True
        out += recs
        carry = block[consumed:]
        if final:
            break
        if consumed == 0 and len(chunk) == 0:
            break
    return out, False


# ---- random dress for the same records --------------------------------------------------------------------------------
def _wrap(rng, s: bytes, eol: bytes) -> bytes:
    if not s or rng.random() < 0.4:
        return s
    width = rng.choice((1, 7, 50, 60, 70, 80))
    return eol.join(s[o : o + width] for o in range(0, len(s), width))


def dress(rng, recs, fmt, style):
    """Bytes of a FASTA / FASTQ file holding `recs` [(id, seq)].  style: dict of switches."""
    eol = b"\r\n" if style.get("crlf") else b"\n"
    out = []
    for rid, s in recs:
        if style.get("lower") and rng.random() < 0.3:
            s = s.lower()
        if fmt == "fasta":
            head = (b";" if style.get("semicolon") and rng.random() < 0.3 else b">") + (b"  " if style.get("id_blanks") and rng.random() < 0.5 else b"") + rid
            body = _wrap(rng, s, eol) if style.get("wrap") else s
            if style.get("blanks") and len(body) > 4 and rng.random() < 0.5:
                p = rng.randrange(1, len(body) - 1)
                body = body[:p] + rng.choice((b" ", b"\t", b" 10 ", eol + eol)) + body[p:]
            out.append(head + eol + body + eol)
            if style.get("blank_lines") and rng.random() < 0.2:
                out.append(eol)
        else:
            seq = _wrap(rng, s, eol) if style.get("wrap") else s
            qual = bytes(rng.choice(b"!#5?IJ~@+>" if style.get("odd_quals") else b"!#5?IJ~") for _ in range(len(s)))
            qual = _wrap(rng, qual, eol) if style.get("wrap") else qual
            plus = b"+" + (rid if style.get("plus_id") and rng.random() < 0.5 else b"")
            if style.get("blanks") and len(seq) > 4 and rng.random() < 0.3:
                p = rng.randrange(1, len(seq) - 1)
                seq = seq[:p] + rng.choice((b" ", b"\t")) + seq[p:]
            out.append(b"@" + rid + eol + seq + eol + plus + eol + qual + eol)
    data = b"".join(out)
    if style.get("no_final_newline"):
        data = data.rstrip(b"\r\n")
    elif style.get("trailing_newlines"):
        data += eol * rng.choice((1, 2))
    if style.get("leading_blank"):
        data = eol + data
    return data


def _lines(path):
    """Sorted lines of an output file, as bytes (ids may hold a carriage return)."""
    if not os.path.exists(path):
        return []
    with open(path, "rb") as f:
        return sorted(l for l in f.read().split(b"\n") if l)


STYLES = ["wrap", "blanks", "blank_lines", "crlf", "lower", "semicolon", "id_blanks", "plus_id", "no_final_newline"]
RARE_STYLES = ["odd_quals", "trailing_newlines", "leading_blank", "empty_seq"]  # drawn with a lower probability


def differential_case(L, seed, tmp, ref_bin):
    """-> (ok, description); ok is None when the reference binary itself crashed on the input.  Reference binary on the
    dressed file vs oracle classification of this reader's records."""
    rng = random.Random(seed)
    k = rng.choice((8, 12, 19))
    w = k + rng.choice((0, 4, 12))
    ibf = os.path.join(tmp, "r%d.ibf" % seed)
    genomes = F.make_db(rng, ibf, k, w)
    recs = F.make_reads(rng, genomes, rng.choice((1, 7, 60)), w)
    fmt = rng.choice(("fasta", "fastq"))
    style = {s: rng.random() < 0.35 for s in STYLES}
    style.update({s: rng.random() < 0.12 for s in RARE_STYLES})
    if style["empty_seq"] and recs:
        j = rng.randrange(len(recs))
        recs[j] = (recs[j][0], b"")
    bad_at = None
    if rng.random() < 0.35 and recs:  # an illegal letter somewhere: parse error
        bad_at = rng.randrange(len(recs))
        rid, s = recs[bad_at]
        p = rng.randrange(max(1, len(s)))
        recs[bad_at] = (rid, s[:p] + rng.choice((b"X", b"E", b"*", b"-", b"Z", b"@")) + s[p + 1 :])
    data = dress(rng, recs, fmt, style)
    path = os.path.join(tmp, "r%d.%s" % (seed, "fa" if fmt == "fasta" else "fq"))
    with open(path, "wb") as f:
        f.write(data)
    n_reads = rng.choice((1, 3, 400))
    cutoff = rng.choice((0.0, 0.2, 0.6))
    out = os.path.join(tmp, "r%d_ref" % seed)
    try:
        pr = subprocess.run([ref_bin, "-r", path, "-i", ibf, "-c", str(cutoff), "-d", "1", "-a", "-u", "-o", out, "-t", "2", "--quiet", "--n-reads", str(n_reads)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=8)
    except subprocess.TimeoutExpired:
        return None, "seed %d: the reference binary hangs on this input (%s, %s)" % (seed, fmt, [s for s in STYLES + RARE_STYLES if style[s]])
    desc = "seed %d %s k=%d w=%d reads=%d n_reads=%d bad_at=%s style=%s" % (seed, fmt, k, w, len(recs), n_reads, bad_at, [s for s in STYLES + RARE_STYLES if style[s]])
    if pr.returncode < 0 or "free():" in pr.stderr or "corrupted" in pr.stderr:
        return None, desc + " reference crashed: " + pr.stderr[-120:].strip()  # seen with CRLF + wrapped FASTA: heap corruption inside the reference
    if pr.returncode != 0:
        return False, desc + " reference failed: " + pr.stderr[-200:]
    want_all, want_unc = _lines(out + ".all"), _lines(out + ".unc")
    mine, _err = read_file(L, data, rng.choice((64, 257, 4096, 1 << 20)), n_reads)
    from ganon_b200 import formats

    filt = O.OracleFilter.from_ibf_file(formats.read_ibf(ibf), cutoff)
    res = O.classify_level([filt], [(i, s, None) for i, s in mine], 1.0, 1.0)
    got_all = sorted(b"%s\t%s\t%d" % (r["id"], t.encode(), c) for r in res for t, c in r["matches"])
    got_unc = sorted(r["id"] for r in res if not r["matches"])
    ok = got_all == want_all and got_unc == want_unc
    if not ok:
        desc += " | ref all=%d unc=%d, mine all=%d unc=%d" % (len(want_all), len(want_unc), len(got_all), len(got_unc))
    return ok, desc


if __name__ == "__main__":
    import sys
    import tempfile

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tmp = tempfile.mkdtemp()
    L = lib(tmp)
    bad = crashed = 0
    for seed in range(first, first + n):
        ok, desc = differential_case(L, seed, tmp, F.REF_BIN)
        if ok is None:
            crashed += 1
            print("SKIP", desc)
        elif not ok:
            bad += 1
            print("MISMATCH", desc)
    print("%d cases, %d mismatches, %d skipped (reference crashed)" % (n, bad, crashed))
