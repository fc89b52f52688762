"""Writes tests/golden/lca_kats.json: the two taxonomies the reference's LCA known-answer test loads
(/root/reference/tests/utils/data/lca/{tree,ncbi}.tax, node <tab> parent) together with the expected values of
tests/utils/LCA.test.cpp:19-115, so that the test can run where /root/reference does not exist (the GPU box).

Run in the build container:  python tests/golden/make_golden_refkats.py
"""
import json
import os

REF = "/root/reference/tests/utils/data/lca"
HERE = os.path.dirname(os.path.abspath(__file__))

# (expected LCA, nodes) -- LCA.test.cpp:28-31 (general), :73-80 (order independence), :100-109 (NCBI).
# The reference fills a map keyed by the expected value, so of several entries with the same key only the last one is ever
# checked: for the toy tree all of them are kept here (they hold), for the NCBI subset the two overwritten entries with key
# "1" are left out -- in that file the LCA of a bacterium and an archaeon is 131567 (cellular organisms), not the root.
CASES = {
    "tree": [
        ["D0", ["E0", "E1"]], ["C3", ["C3", "F4"]], ["A0", ["G0", "C3", "D5"]], ["1", ["G0", "G5"]],
        ["B1", ["B1", "C2"]], ["B1", ["C2", "B1"]],
        ["B0", ["C0", "E1", "F2"]], ["B0", ["C0", "F2", "E1"]], ["B0", ["F2", "C0", "E1"]], ["B0", ["F2", "E1", "C0"]], ["B0", ["E1", "F2", "C0"]], ["B0", ["E1", "C0", "F2"]],
    ],
    "ncbi": [
        ["1224", ["366602", "470"]], ["2", ["366602", "470", "1406"]], ["2290931", ["2223", "51589"]], ["10239", ["2025595", "491893"]],
        ["1", ["366602", "470", "1406", "2223", "51589", "2025595", "491893"]],
    ],
}

out = {}
for name, cases in CASES.items():
    edges = []
    with open(os.path.join(REF, name + ".tax")) as f:
        for line in f:
            a = line.rstrip("\n").split("\t")
            if len(a) >= 2:
                edges.append([a[0], a[1]])
    out[name] = {"edges": edges, "cases": cases}
with open(os.path.join(HERE, "lca_kats.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print({k: (len(v["edges"]), len(v["cases"])) for k, v in out.items()})

# ---- tests/ganon-build/GanonBuild.test.cpp: the literal sequences its sections build filters from (:118-128 `seqs`,
# :521-532 `seqs2` for --min-length, :562-570 `seqs3` = the ones of seqs2 that are at least 50 bp long)
import re

src = open("/root/reference/tests/ganon-build/GanonBuild.test.cpp").read()
seqsets = {}
for m in re.finditer(r"aux::sequences_type\s+(\w+)\s*\{(.*?)\};", src, re.S):
    seqsets[m.group(1)] = re.findall(r'"([ACGT]+)"_dna4', m.group(2))
assert [len(seqsets[k]) for k in ("seqs", "seqs2", "seqs3")] == [10, 10, 7]
with open(os.path.join(HERE, "build_kats.json"), "w") as f:
    json.dump(seqsets, f, separators=(",", ":"))
print({k: len(v) for k, v in seqsets.items()})
