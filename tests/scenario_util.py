"""Helpers shared by the parity tests: parse the golden scenarios' ganon-classify argv."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_scenarios():
    with open(os.path.join(GOLDEN, "scenarios.json")) as f:
        return json.load(f)


def load_hibf_scenarios():
    with open(os.path.join(GOLDEN, "scenarios_hibf.json")) as f:
        return json.load(f)


def expand(args, dbs):
    """Fill the {golden}/{tmp} placeholders; {tmp}/X.ibf -> decompressed fixture path."""
    out = []
    for a in args:
        a = a.replace("{golden}", GOLDEN)
        for n, p in dbs.items():
            a = a.replace("{tmp}/%s.ibf" % n, p)
        if "synth_hibf" in dbs:
            a = a.replace("{tmp}/synth.hibf", dbs["synth_hibf"])
        out.append(a)
    return out


def parse_args(argv):
    """Tiny re-implementation of the ganon-classify flag grammar for the oracle-level tests
    (CommandLineParser.cpp:15-45; broadcast rules Config.hpp:175-245)."""
    cfg = dict(single=[], paired=[], ibf=[], tax=[], labels=["H1"], rel_cutoff=[0.2], rel_filter=[0.0], fpr_query=[1.0], flags=set())
    names = {"-r": "single", "-p": "paired", "-i": "ibf", "-x": "tax", "-y": "labels", "-c": "rel_cutoff", "-d": "rel_filter", "-f": "fpr_query"}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in names:
            v = argv[i + 1].split(",")
            if names[a] in ("rel_cutoff", "rel_filter", "fpr_query"):
                v = [float(x) for x in v]
            cfg[names[a]] = v
            i += 2
        elif a == "--hibf":
            cfg["flags"].add(a)
            i += 1
        else:
            cfg["flags"].add(a)
            i += 1
    n = len(cfg["ibf"])
    if len(cfg["labels"]) == 1:
        cfg["labels"] = cfg["labels"] * n
    if len(cfg["rel_cutoff"]) == 1:
        cfg["rel_cutoff"] = cfg["rel_cutoff"] * n
    uniq = sorted(set(cfg["labels"]))
    if len(cfg["rel_filter"]) == 1:
        cfg["rel_filter"] = cfg["rel_filter"] * len(uniq)
    if len(cfg["fpr_query"]) == 1:
        cfg["fpr_query"] = cfg["fpr_query"] * len(uniq)
    # parse_hierarchy GanonClassify.cpp:353-401: rel_filter/fpr_query are assigned in order of FIRST APPEARANCE
    # of each label in --hierarchy-labels, levels are then processed in sorted label order
    levels = {}
    order = []
    for j, lab in enumerate(cfg["labels"]):
        if lab not in levels:
            levels[lab] = dict(filters=[], rel_filter=cfg["rel_filter"][len(order)], fpr_query=cfg["fpr_query"][len(order)])
            order.append(lab)
        levels[lab]["filters"].append((cfg["ibf"][j], cfg["rel_cutoff"][j]))
    cfg["levels"] = [(lab, levels[lab]) for lab in sorted(levels)]
    return cfg


def expected_lines(name, ext):
    p = os.path.join(GOLDEN, "expected", name + "." + ext)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return sorted(l.rstrip("\n") for l in f)
