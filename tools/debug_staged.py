"""Scratch: why does stage + run_staged + finish_staged differ from classify on the FASTA fixture?"""
import gzip, os, shutil, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganon_b200.classify import Database, Session, result_text
G = "tests/golden"
d = tempfile.mkdtemp()
p = os.path.join(d, "synth.ibf")
with gzip.open(os.path.join(G, "synth.ibf.gz"), "rb") as fi, open(p, "wb") as fo:
    shutil.copyfileobj(fi, fo)
db = Database.open(p)
for name in ("reads.fa", "reads.se.fq"):
    fq = open(os.path.join(G, name), "rb").read()
    for env in ("", "host"):
        os.environ["GANON_B200_HOST_INDEX"] = "1" if env else "0"
        mk = lambda: Session([db], [0.1], [0.5], [1.0], output_all=True, output_unclassified=True)
        s = mk(); r = s.classify(fq, final=True)
        a = set(result_text(r, "all").decode().splitlines()); ua = set(result_text(r, "unc").decode().splitlines()); ra = s.report()
        s = mk(); s.stage(fq, final=True); t = s.run_staged(); r = s.finish_staged()
        b = set(result_text(r, "all").decode().splitlines()); ub = set(result_text(r, "unc").decode().splitlines()); rb = s.report()
        s = mk(); s.stage(fq, final=True); r = s.finish_staged()
        c = set(result_text(r, "all").decode().splitlines())
        print(name, env, "classify-only:", sorted(a - b)[:5], "staged-only:", sorted(b - a)[:5], "unc", sorted(ua ^ ub)[:5], "rep equal", ra == rb, "stage+finish == classify", a == c, "n", r.n_reads, t.n_reads)
