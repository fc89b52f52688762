"""`ganon-classify` command line: the flag grammar of the reference binary (CommandLineParser.cpp:15-45, cxxopts:
short/long names, comma-separated lists, boolean switches, `--opt=value`), exit codes of main.cpp:9-16."""
from __future__ import annotations

import sys
from typing import List, Optional

from .classify import VERSION, GanonClassifyConfig, run

# long name -> (short, kind, config attribute)
_OPTS = {
    "single-reads": ("r", "strs", "single_reads"),
    "paired-reads": ("p", "strs", "paired_reads"),
    "batch-reads": ("b", "strs", "batch_reads"),
    "ibf": ("i", "strs", "ibf"),
    "tax": ("x", "strs", "tax"),
    "hierarchy-labels": ("y", "strs", "hierarchy_labels"),
    "rel-cutoff": ("c", "floats", "rel_cutoff"),
    "rel-filter": ("d", "floats", "rel_filter"),
    "fpr-query": ("f", "floats", "fpr_query"),
    "output-prefix": ("o", "str", "output_prefix"),
    "output-lca": ("l", "bool", "output_lca"),
    "output-all": ("a", "bool", "output_all"),
    "output-unclassified": ("u", "bool", "output_unclassified"),
    "output-stats": ("z", "bool", "output_stats"),
    "output-single": ("s", "bool", "output_single"),
    "hibf": (None, "bool", "hibf"),
    "skip-lca": (None, "bool", "skip_lca"),
    "tax-root-node": (None, "str", "tax_root_node"),
    "threads": ("t", "int", "threads"),
    "n-batches": (None, "int", "n_batches"),
    "n-reads": (None, "int", "n_reads"),
    "verbose": (None, "bool", "verbose"),
    "quiet": (None, "bool", "quiet"),
    "device": (None, "int", "device"),  # extension: GPU ordinal
    "devices": (None, "ints", "devices"),  # extension: several GPUs, the databases bin-sharded over them
    "hbm-budget-gb": (None, "float", "hbm_budget_gb"),  # extension: host-resident tier for filters above this many GiB
    # extension: `ganon classify --multiple-matches em` without the round trip through the .all file (src/ganon/reassign.py)
    "reassign-em": (None, "bool", "reassign_em"),
    "em-max-iter": (None, "int", "em_max_iter"),
    "em-threshold": (None, "floats", "em_threshold"),
    "help": ("h", "flag", None),
    "version": ("v", "flag", None),
}
_SHORT = {v[0]: k for k, v in _OPTS.items() if v[0]}

HELP = """Ganon classifier (B200)
Usage:
  ganon-classify [OPTION...]

  -r, --single-reads arg      single-end reads file[s] (comma-separated, flat or gzipped)
  -p, --paired-reads arg      paired-end reads file[s] (comma-separated, flat or gzipped)
  -b, --batch-reads arg       file describing several files of single- or paired-end reads: prefix <tab> file1 [<tab> file2]
  -i, --ibf arg               ibf file[s] from ganon-build (comma-separated)
  -x, --tax arg               tax file[s] from ganon-build for LCA calculation (comma-separated)
  -y, --hierarchy-labels arg  Hierarchy labels to define level for classification. Default: H1
  -c, --rel-cutoff arg        Relative cutoff. One or one per filter (comma-separated). Default: 0.2
  -d, --rel-filter arg        Relative filter. one or one per hierarchy label (comma-separated). Default: 0.0
  -f, --fpr-query arg         Min. False positive for a query. one or one per hierarchy label. Default: 1.0
  -o, --output-prefix arg     Output prefix (prefix.rep, [prefix.one, prefix.all, prefix.unc])
  -l, --output-lca            Runs and outputs file with lca classification (prefix.one)
  -a, --output-all            Outputs file with all matches (prefix.all)
  -u, --output-unclassified   Outputs unclassified read ids (prefix.unc)
  -z, --output-stats          Outputs classification statistics (prefix.sta)
  -s, --output-single         Do not split output files (one and all) with multi-level --hierarchy-labels
      --hibf                  Input is an Hierarchical IBF (.hibf) generated from raptor.
      --skip-lca              Skip LCA step.
      --tax-root-node arg     Define alternative root node for LCA. Default: 1
  -t, --threads arg           Number of host threads for the finishing stage
      --n-batches arg         (accepted for compatibility)
      --n-reads arg           Number of reads for each batch. Default: 400
      --device arg            CUDA device ordinal. Default: 0
      --devices arg           Several CUDA devices (comma-separated): every .ibf is split by bin columns over them
                              (databases larger than one GPU's memory; one process per GPU, results identical)
      --hbm-budget-gb arg     Keep at most this many GiB of a flat .ibf in GPU memory; the rest stays in host memory
                              and is streamed per batch (host-resident tier; results identical)
      --reassign-em           EM reassignment of reads with several matches from the matches kept on the GPU
                              (what `ganon classify --multiple-matches em` does from the .all file): writes prefix.one
                              and the reassigned prefix.rep
      --em-max-iter arg       Max. number of EM iterations, 0 = until convergence. Default: 10
      --em-threshold arg      Convergence threshold of the EM. Default: 0
      --verbose               Verbose output mode
      --quiet                 Quiet output mode
  -h, --help                  Print help
  -v, --version               Show version
"""


class CliError(Exception):
    pass


def parse(argv: List[str]) -> Optional[GanonClassifyConfig]:
    """Returns None for -h / -v / no arguments (the caller maps that to the reference's exit code)."""
    if len(argv) == 0:
        print("Try 'ganon-classify -h/--help' for more information.", file=sys.stderr)
        return None
    cfg = GanonClassifyConfig()
    seen = {}
    flags = set()
    i = 0

    def take(name: str, inline: Optional[str]) -> None:
        nonlocal i
        _short, kind, attr = _OPTS[name]
        if kind == "flag":
            flags.add(name)
            return
        if kind == "bool":
            val = True if inline is None else inline.lower() not in ("0", "false", "f")
            setattr(cfg, attr, val)
            return
        if inline is None:
            i += 1
            if i >= len(argv):
                raise CliError("Option '%s' is missing an argument" % name)
            inline = argv[i]
        try:
            if kind == "strs":
                vals = inline.split(",")
                if name in seen:
                    vals = getattr(cfg, attr) + vals
                setattr(cfg, attr, vals)
            elif kind == "floats":
                vals = [float(x) for x in inline.split(",")]
                if name in seen:
                    vals = getattr(cfg, attr) + vals
                setattr(cfg, attr, vals)
            elif kind == "ints":
                setattr(cfg, attr, [int(x) for x in inline.split(",")])
            elif kind == "float":
                setattr(cfg, attr, float(inline))
            elif kind == "int":
                setattr(cfg, attr, int(inline))
            else:
                setattr(cfg, attr, inline)
        except ValueError:
            raise CliError("Argument '%s' failed to parse" % inline)
        seen[name] = True

    while i < len(argv):
        a = argv[i]
        if a.startswith("--"):
            name, _, inline = a[2:].partition("=")
            if name not in _OPTS:
                raise CliError("Option '%s' does not exist" % name)
            take(name, inline if "=" in a else None)
        elif a.startswith("-") and len(a) > 1:
            body = a[1:]
            j = 0
            while j < len(body):
                ch = body[j]
                if ch not in _SHORT:
                    raise CliError("Option '%s' does not exist" % ch)
                name = _SHORT[ch]
                if _OPTS[name][1] in ("bool", "flag"):
                    take(name, None)
                    j += 1
                else:
                    rest = body[j + 1 :]
                    take(name, rest if rest else None)
                    break
        else:
            raise CliError("Option '%s' does not exist" % a)
        i += 1
    if "help" in flags:
        print(HELP, file=sys.stderr)
        return None
    if "version" in flags:
        print("version: " + VERSION, file=sys.stderr)
        return None
    return cfg


def main(argv: Optional[List[str]] = None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    try:
        cfg = parse(argv)
    except CliError as e:
        print(str(e), file=sys.stderr)
        return 1
    if cfg is None:
        return 1 if len(argv) == 0 else 0  # main.cpp:14-16
    return 0 if run(cfg) else 1


if __name__ == "__main__":
    sys.exit(main())
