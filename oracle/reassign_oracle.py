"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ganon's EM reassignment of multi-matching reads
(`/root/reference/src/ganon/reassign.py`, run by `ganon classify --multiple-matches em` after the binary,
`src/ganon/classify.py:76-88`).  Nothing here is on the product path; only tests/ may import it.

Pinned: tests/test_reassign_cpu.py checks it against tests/golden/expected_em/, which tests/golden/make_golden_em.py
produced by running the unmodified reference module on the reference binary's own `.all` / `.rep` outputs.

The restatement works on arrays instead of the reference's dictionaries:
  * targets are numbered in order of first appearance in the `.all` file (reassign.py:75, 84-88: auto-increment dict),
    reads likewise by id (a repeated id extends the earlier read, as the dict of lists does);
  * a read with one match adds 1 to its target's initial weight (reassign.py:94-101); `prob = weight / total`;
  * every iteration (reassign.py:110-141) gives each multi-matching read to its match with the highest probability --
    the first one wins ties, and a probability of 0 never wins against the first match (get_top_match, 229-241) --
    then recomputes the probabilities from the redistributed counts and sums |old - new| in target order;
  * stop when diff <= threshold or after max_iter iterations (max_iter = 0: until convergence), 136-139;
  * `.one`: one line per read in order of first appearance, the unique match or the top match under the FINAL
    probabilities (149-179);  new `.rep`: the old lines of targets present in the `.all`, lca column = reassigned - unique
    with the counts of the LAST iteration (188-214), then the '#' lines.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def _top(matches: Sequence[Tuple[int, int]], prob: Sequence[float]) -> Tuple[int, int]:
    best_t, best_k = matches[0]
    best_p = 0
    for t, k in matches:
        if prob[t] > best_p:
            best_p, best_t, best_k = prob[t], t, k
    return best_t, best_k


def em_all_file(all_text: str, threshold: float, max_iter: int):
    """One `.all` file -> (target names, reassigned counts of the last iteration, `.one` text, iterations run)."""
    target_id: Dict[str, int] = {}
    names: List[str] = []
    read_id: Dict[str, int] = {}
    read_names: List[str] = []
    reads: List[List[Tuple[int, int]]] = []
    for line in all_text.splitlines():
        rid, target, k = line.rstrip().split("\t")
        if target not in target_id:
            target_id[target] = len(names)
            names.append(target)
        if rid not in read_id:
            read_id[rid] = len(reads)
            read_names.append(rid)
            reads.append([])
        reads[read_id[rid]].append((target_id[target], int(k)))
    n_t = len(names)
    initial = [0] * n_t
    uniq_reads = 0
    for m in reads:
        if len(m) == 1:
            initial[m[0][0]] += 1
            uniq_reads += 1
    total = len(reads)
    denom0 = uniq_reads if uniq_reads else 1
    prob = [w / denom0 for w in initial]
    it = 0
    while True:
        counts = list(initial)
        for m in reads:
            if len(m) > 1:
                counts[_top(m, prob)[0]] += 1
        diff = 0
        for t in range(n_t):
            p = counts[t] / total
            diff += abs(prob[t] - p)
            prob[t] = p
        if diff <= threshold:
            break
        if max_iter > 0 and it == max_iter - 1:
            break
        it += 1
    out = []
    for rid, m in zip(read_names, reads):
        t, k = m[0] if len(m) == 1 else _top(m, prob)
        out.append("%s\t%s\t%d\n" % (rid, names[t], k))
    return names, counts, "".join(out), it + 1


def reassign_texts(rep_text: str, all_texts: Dict[str, str], threshold: float = 0, max_iter: int = 10):
    """`all_texts`: hierarchy label -> `.all` text in `.rep` order of first appearance ("" = the single `.all` of a run
    with one label or --output-single).  Returns ({label: `.one` text}, new `.rep` text)."""
    info = [l.rstrip() for l in rep_text.splitlines() if l.startswith("#")]
    rows = [l.rstrip().split("\t") for l in rep_text.splitlines() if l and not l.startswith("#")]
    ones: Dict[str, str] = {}
    new_rep: List[str] = []
    for label, text in all_texts.items():
        names, counts, one, _ = em_all_file(text, threshold, max_iter)
        ones[label] = one
        tid = {n: i for i, n in enumerate(names)}
        for f in rows:
            if (label == "" or f[0] == label) and f[1] in tid:
                rank = f[5] if len(f) >= 6 else ""
                name = f[6] if len(f) >= 7 else ""
                new_rep.append("\t".join([f[0], f[1], f[2], str(int(f[3])), str(counts[tid[f[1]]] - int(f[3])), rank, name]))
    return ones, "".join(l + "\n" for l in new_rep + info)


def all_files_of(rep_text: str, available: Sequence[str]) -> List[str]:
    """Which `.all` files a `.rep` refers to (reassign.py:37-60): one per hierarchy label if `<prefix>.<label>.all`
    exists, else the single `<prefix>.all`.  `available`: the suffixes that exist ("all", "<label>.all")."""
    labels: List[str] = []
    for l in rep_text.splitlines():
        if l and not l.startswith("#"):
            h = l.split("\t")[0]
            if h not in labels:
                labels.append(h)
    out: List[str] = []
    for h in labels:
        if h + ".all" in available:
            out.append(h)
        elif "all" in available:
            return [""]
    return out
