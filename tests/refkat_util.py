"""The known-answer tests of the reference's own test suite for the classify path, restated as data:
/root/reference/tests/ganon-classify/GanonClassify.test.cpp (two SCENARIOs: "classifying reads without errors" :186-797 and
"classifying reads with errors" :799-1228).  The reference builds tiny filters from literal sequences with GanonBuild::run,
classifies literal reads with GanonClassify::run and asserts exact per-read / per-target counts; the same inputs are written
here (one FASTA file per sequence, as aux::SeqTarget does, tests/aux/Aux.hpp:142-237) and every case carries the numbers the
reference asserts (file:line in `ref`).

Consumers: tests/test_reference_kats.py -- oracle (CPU), the unmodified reference binaries where oracle/_ref exists (CPU),
and the drop-in command line on the GPU.
"""
import os

# a dna4 literal turns every character that is not ACGT into 'A' (the reference's comment at :813)
_L = lambda s: s.replace("-", "A")

READS = {
    # scenario 1 (:192-199), 14 bp
    "readA": "A" * 14, "readC": "C" * 14, "readT": "T" * 14, "readG": "G" * 14,
    # :635-637
    "readCG": "CG" * 7,
    # scenario 2 (:804-811), 12 bp; :1193-1195 with one error
    "readF": "CTCGTGTTTCCT", "readR": "ACCAAGAGGCCC",
    "readFe1": _L("CTCGTGTTTCC-"), "readRe1": _L("ACCAAGAGGCC-"),
}

# build name -> (k, w, [(sequence header, target, sequence)])   max_fp 0.01 everywhere
_REFS1 = [("seqA", "A", "A" * 20), ("seqC", "C", "C" * 20), ("seqT", "T", "T" * 20), ("seqG", "G", "G" * 20)]  # :202-210
_REFS2 = [("seqA2", "A2", "A" * 20), ("seqCG", "CG", "CG" * 10)]  # :640-646
_REFS3 = [(t, t, _L(s)) for t, s in [  # :833-842 (the seventh name has no sequence)
    ("e0", "CTCGTGTTTCCT----GGGCCTCTTGGT"), ("e1F", "CTC-TGTTTCCT----GGGCCTCTTGGT"), ("e1F_e1R", "CTC-TGTTTCCT----GGG-CTCTTGGT"),
    ("e1F_e2R", "CTC-TGTTTCCT----GGG-CTCT-GGT"), ("e2F_e1R", "CTC-TGTT-CCT----GGG-CTCTTGGT"), ("e2F_e2R", "CTC-TGTT-CCT----GGG-CTCT-GGT")]]
BUILDS = {
    "b1": (10, 10, _REFS1),    # :214-222
    "b1ws": (10, 12, _REFS1),  # :586-595
    "b2": (10, 10, _REFS2),    # :652-660
    "b3": (4, 4, _REFS3),      # :846-854
    "b3ws": (4, 6, _REFS3),    # :1106-1115
}

_TAX1 = {"A": "AT", "C": "CG", "T": "AT", "G": "CG", "CG": "ATCG", "AT": "ATCG", "ATCG": "1"}  # :523-530
TAXES = {
    "tax1": _TAX1,
    "tax1_noA": {k: v for k, v in _TAX1.items() if k != "A"},  # :560-566
    "tax2": {"A2": "AT", "CG": "ATCG", "AT": "ATCG", "ATCG": "1"},  # :672-675
}


def _c(name, ref, ibf, single=(), paired=(), cutoff=0.0, relfilter=1.0, fpr=None, tax=(), labels=(), flags=(), all_=None, one=None, absent=(), present=()):
    return dict(name=name, ref=ref, ibf=list(ibf), single=list(single), paired=list(paired), cutoff=cutoff, relfilter=relfilter, fpr=fpr, tax=list(tax),
                labels=list(labels), flags=list(flags), all=all_ or {}, one=one, absent=list(absent), present=list(present))


_DEF = ("--output-all", "--output-lca", "--output-unclassified", "--output-stats")  # defaultConfig :21-33
_E = "GanonClassify.test.cpp"
CASES = [
    # ---- without errors
    _c("single", _E + ":253-269", ["b1"], single=["readA"], flags=_DEF, all_={"readA": {"A": 5, "T": 5}}),
    _c("single_wo_lca", _E + ":271-284", ["b1"], single=["readA"], flags=("--output-all", "--output-unclassified", "--output-stats"), all_={"readA": {"A": 5, "T": 5}}, absent=[".one"]),
    _c("single_wo_all", _E + ":286-299", ["b1"], single=["readA"], flags=("--output-lca", "--output-unclassified", "--output-stats"), absent=[".all"]),
    _c("single_wo_stats", _E + ":301-314", ["b1"], single=["readA"], flags=("--output-all", "--output-lca", "--output-unclassified"), all_={"readA": {"A": 5, "T": 5}}, absent=[".sta"]),
    _c("paired", _E + ":319-336", ["b1"], paired=["readA", "readT"], flags=_DEF, all_={"readA": {"A": 10, "T": 10}}),
    _c("single_and_paired", _E + ":338-362", ["b1"], single=["readC", "readG"], paired=["readA", "readT"], flags=_DEF,
       all_={"readA": {"A": 10, "T": 10}, "readC": {"C": 5, "G": 5}, "readG": {"C": 5, "G": 5}}),
    _c("tax", _E + ":510-546", ["b1"], single=["readA"], tax=["tax1"], flags=_DEF, all_={"readA": {"A": 5, "T": 5}}, one={"readA": {"AT": 5}}),
    _c("incomplete_tax", _E + ":548-582", ["b1"], single=["readA"], tax=["tax1_noA"], flags=_DEF, all_={"readA": {"A": 5, "T": 5}}, one={"readA": {"1": 5}}),
    _c("window_size", _E + ":584-610", ["b1ws"], single=["readA"], flags=_DEF, all_={"readA": {"A": 1, "T": 1}}),
    _c("window_size_paired", _E + ":612-628", ["b1ws"], paired=["readA", "readT"], flags=_DEF, all_={"readA": {"A": 2, "T": 2}}),
    _c("two_ibf_one_level", _E + ":678-697", ["b1", "b2"], single=["readA", "readCG"], flags=_DEF, all_={"readA": {"A": 5, "T": 5, "A2": 5}, "readCG": {"CG": 5}}),
    _c("two_ibf_one_level_tax", _E + ":699-724", ["b1", "b2"], single=["readA", "readCG"], tax=["tax1", "tax2"], flags=_DEF,
       all_={"readA": {"A": 5, "T": 5, "A2": 5}, "readCG": {"CG": 5}}, one={"readA": {"AT": 5}, "readCG": {"CG": 5}}),
    _c("two_levels", _E + ":727-748", ["b1", "b2"], single=["readA", "readCG"], labels=["one", "two"], flags=_DEF + ("--output-single",),
       all_={"readA": {"A": 5, "T": 5}, "readCG": {"CG": 5}}),
    _c("two_levels_tax", _E + ":750-776", ["b1", "b2"], single=["readA", "readCG"], labels=["one", "two"], tax=["tax1", "tax2"], flags=_DEF + ("--output-single",),
       all_={"readA": {"A": 5, "T": 5}, "readCG": {"CG": 5}}, one={"readA": {"AT": 5}, "readCG": {"CG": 5}}),
    _c("two_levels_split_files", _E + ":778-794", ["b1", "b2"], single=["readA", "readCG"], labels=["one", "two"], flags=_DEF, present=[".one.all", ".two.all"]),
    # ---- with errors (k = w = 4; expected maximum 4-mer counts in the table at :822-831)
    _c("c045_f0", _E + ":856-870", ["b3"], single=["readF"], cutoff=0.45, relfilter=0.0, flags=_DEF, all_={"readF": {"e0": 9}}),
    _c("c045_f0_paired", _E + ":872-885", ["b3"], paired=["readF", "readR"], cutoff=0.45, relfilter=0.0, flags=_DEF, all_={"readF": {"e0": 18}}),
    _c("c02_f08", _E + ":888-905", ["b3"], single=["readF"], cutoff=0.2, relfilter=0.8, flags=_DEF, all_={"readF": {"e0": 9, "e1F": 5, "e1F_e1R": 5, "e1F_e2R": 5}}),
    _c("c02_f08_paired", _E + ":907-922", ["b3"], paired=["readF", "readR"], cutoff=0.2, relfilter=0.8, flags=_DEF, all_={"readF": {"e0": 18, "e1F": 14, "e1F_e1R": 10}}),
    _c("c02_f08_q", _E + ":925-940", ["b3"], single=["readF"], cutoff=0.2, relfilter=0.8, fpr=1e-10, flags=_DEF, all_={"readF": {"e0": 9, "e1F_e2R": 5}}),
    _c("c02_f08_q_paired", _E + ":942-958", ["b3"], paired=["readF", "readR"], cutoff=0.2, relfilter=0.8, fpr=1e-10, flags=_DEF, all_={"readF": {"e0": 18, "e1F": 14, "e1F_e1R": 10}}),
    _c("c06_f1", _E + ":961-975", ["b3"], single=["readF"], cutoff=0.6, relfilter=1.0, flags=_DEF, all_={"readF": {"e0": 9}}),
    _c("c06_f1_paired", _E + ":977-991", ["b3"], paired=["readF", "readR"], cutoff=0.6, relfilter=1.0, flags=_DEF, all_={"readF": {"e0": 18, "e1F": 14}}),
    _c("c0_f03", _E + ":994-1008", ["b3"], single=["readF"], cutoff=0.0, relfilter=0.3, flags=_DEF, all_={"readF": {"e0": 9}}),
    _c("c0_f03_paired", _E + ":1010-1024", ["b3"], paired=["readF", "readR"], cutoff=0.0, relfilter=0.3, flags=_DEF, all_={"readF": {"e0": 18, "e1F": 14}}),
    _c("c0_f1", _E + ":1027-1045", ["b3"], single=["readF"], flags=_DEF, all_={"readF": {"e0": 9, "e1F": 5, "e1F_e1R": 5, "e1F_e2R": 5, "e2F_e1R": 1, "e2F_e2R": 1}}),
    _c("c0_f1_paired", _E + ":1047-1065", ["b3"], paired=["readF", "readR"], flags=_DEF,
       all_={"readF": {"e0": 18, "e1F": 14, "e1F_e1R": 10, "e1F_e2R": 6, "e2F_e1R": 6, "e2F_e2R": 2}}),
    _c("c0_f1_q", _E + ":1068-1084", ["b3"], single=["readF"], fpr=1e-10, flags=_DEF, all_={"readF": {"e0": 9, "e1F_e2R": 5}}),
    _c("c0_f1_q_paired", _E + ":1086-1101", ["b3"], paired=["readF", "readR"], fpr=1e-10, flags=_DEF, all_={"readF": {"e0": 18, "e1F": 14, "e1F_e1R": 10}}),
    _c("ws6_c1_f0", _E + ":1117-1132", ["b3ws"], single=["readF"], cutoff=1.0, relfilter=0.0, flags=_DEF, all_={"readF": {"e0": 4}}),
    _c("ws6_c1_f0_paired", _E + ":1134-1147", ["b3ws"], paired=["readF", "readR"], cutoff=1.0, relfilter=0.0, flags=_DEF, all_={"readF": {"e0": 8}}),
    _c("ws6_c0_f1", _E + ":1150-1167", ["b3ws"], single=["readF"], flags=_DEF, all_={"readF": {"e0": 4, "e1F": 2, "e1F_e1R": 2, "e1F_e2R": 2}}),
    _c("ws6_c0_f1_paired", _E + ":1169-1185", ["b3ws"], paired=["readF", "readR"], flags=_DEF, all_={"readF": {"e0": 8, "e1F": 6, "e1F_e1R": 4, "e1F_e2R": 2, "e2F_e1R": 2}}),
    _c("read_errors_c07_f0", _E + ":1189-1211", ["b3"], single=["readFe1"], cutoff=0.7, relfilter=0.0, flags=_DEF, all_={"readFe1": {"e0": 8}}),
    _c("read_errors_c07_f0_paired", _E + ":1213-1226", ["b3"], paired=["readFe1", "readRe1"], cutoff=0.7, relfilter=0.0, flags=_DEF, all_={"readFe1": {"e0": 16}}),
]


def write_inputs(d):
    """Reads (one FASTA per read), reference sequences + the builders' input tables, taxonomies (write_tax :170-182)."""
    paths = {"reads": {}, "tsv": {}, "tax": {}}
    for name, seq in READS.items():
        p = os.path.join(d, name + ".fasta")
        with open(p, "w") as f:
            f.write(">%s\n%s\n" % (name, seq))
        paths["reads"][name] = p
    for b, (_k, _w, refs) in BUILDS.items():
        tsv = os.path.join(d, b + ".tsv")
        with open(tsv, "w") as t:
            for header, target, seq in refs:
                p = os.path.join(d, "ref_%s.fasta" % header)
                with open(p, "w") as f:
                    f.write(">%s\n%s\n" % (header, seq))
                t.write("%s\t%s\n" % (p, target))
        paths["tsv"][b] = tsv
    for name, tax in TAXES.items():
        p = os.path.join(d, name + ".tax")
        with open(p, "w") as f:
            f.write("1\t0\troot\troot\n")
            for node, parent in sorted(tax.items()):
                f.write("%s\t%s\trank-%s\tname-%s\n" % (node, parent, node, node))
        paths["tax"][name] = p
    return paths


def argv_of(case, paths, ibf_paths):
    """ganon-classify arguments of a case (without -o / -t / --quiet)."""
    a = ["-i", ",".join(ibf_paths[b] for b in case["ibf"])]
    if case["single"]:
        a += ["-r", ",".join(paths["reads"][r] for r in case["single"])]
    if case["paired"]:
        a += ["-p", ",".join(paths["reads"][r] for r in case["paired"])]
    if case["tax"]:
        a += ["-x", ",".join(paths["tax"][t] for t in case["tax"])]
    if case["labels"]:
        a += ["-y", ",".join(case["labels"])]
    a += ["-c", repr(case["cutoff"]), "-d", repr(case["relfilter"])]
    if case["fpr"] is not None:
        a += ["-f", repr(case["fpr"])]
    return a + list(case["flags"])


def parse_matches(path):
    """`readid <tab> target <tab> count` -> {read: {target: count}} (Res::parse_all_lca :82-101)."""
    out = {}
    with open(path) as f:
        for line in f:
            rid, target, c = line.rstrip("\n").split("\t")
            out.setdefault(rid, {})[target] = int(c)
    return out


def sanity_check(prefix, flags, has_tax):
    """config_classify::sanity_check :147-168 on the files of one run (single prefix, files not split by level)."""
    classified = unclassified = matches = 0
    with open(prefix + ".rep") as f:
        for line in f:
            a = line.rstrip("\n").split("\t")
            if a[0] == "#total_classified":
                classified = int(a[1])
            elif a[0] == "#total_unclassified":
                unclassified = int(a[1])
            else:
                matches += int(a[2])
    if "--output-all" in flags and os.path.exists(prefix + ".all"):
        lines = open(prefix + ".all").read().splitlines()
        assert len({l.split("\t")[0] for l in lines}) == classified and len(lines) == matches
    if "--output-lca" in flags and has_tax and os.path.exists(prefix + ".one"):
        lines = open(prefix + ".one").read().splitlines()
        assert len(lines) == classified and len({l.split("\t")[0] for l in lines}) == classified
    if "--output-unclassified" in flags:
        assert len(open(prefix + ".unc").read().splitlines()) == unclassified
    return classified, unclassified
