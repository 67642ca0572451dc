"""GPU: UMI collapse against a host restatement of UMI-tools' published clustering rules.

The reference has no UMI collapse of its own (its umi/ package is an unfinished sketch and its README only times `a ^ b`
against UMI-tools' edit_distance), so there is no reference output to pin this against: "parity unpinned" for N3.  The
checker below restates umi_tools/network.py (UMIClusterer: _get_adj_list_directional / _get_adj_list_cluster,
_get_connected_components_adjacency, _group_directional / _group_cluster)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hamming(a, b):
    return sum(x != y for x, y in zip(a, b))


def umi_tools_groups(umis, counts, threshold, method):
    """-> list of groups (lists of indices), UMI-tools order: components by decreasing count of their seed."""
    n = len(umis)
    adj = {i: [] for i in range(n)}
    for a, b in itertools.combinations(range(n), 2):
        if len(umis[a]) != len(umis[b]) or _hamming(umis[a], umis[b]) > threshold:
            continue
        if method == "cluster":
            adj[a].append(b)
            adj[b].append(a)
        else:
            if counts[a] >= 2 * counts[b] - 1:
                adj[a].append(b)
            if counts[b] >= 2 * counts[a] - 1:
                adj[b].append(a)
    found, comps = set(), []
    for node in sorted(range(n), key=lambda x: counts[x], reverse=True):      # stable: ties in input order
        if node in found:
            continue
        seen, queue = {node}, [node]
        while queue:
            v = queue.pop(0)
            for u in adj[v]:
                if u not in seen:
                    seen.add(u)
                    queue.append(u)
        found.update(seen)
        comps.append(seen)
    observed, groups = set(), []
    for comp in comps:
        g = [v for v in sorted(comp, key=lambda x: (-counts[x], x)) if v not in observed]
        observed.update(g)
        groups.append(g)
    return groups


@pytest.mark.parametrize("method", ["directional", "cluster"])
@pytest.mark.parametrize("threshold", [1, 2])
def test_umi_collapse_matches_umi_tools_rules(sq, method, threshold):
    rng = np.random.default_rng(7 + threshold)
    group_sizes = [1, 2, 37, 300, 0, 900, 32, 33, 31, 5, 17, 24, 32, 8]      # <= 32: a warp per group; larger: a CTA per group
    umis, counts, goff = [], [], [0]
    for gs in group_sizes:
        seeds = ["".join(rng.choice(list("ACGT"), size=10)) for _ in range(max(1, gs // 6))]
        seen = set()
        while len(seen) < gs:                             # true UMIs plus 1-2 mismatch errors of them, all distinct
            s = list(seeds[rng.integers(len(seeds))])
            for _ in range(rng.integers(0, 3)):
                s[rng.integers(10)] = "ACGT"[rng.integers(4)]
            seen.add("".join(s))
        g = sorted(seen)
        rng.shuffle(g)
        umis += g
        counts += [int(x) for x in np.maximum(1, rng.geometric(0.15, size=len(g)) * rng.choice([1, 1, 1, 20], size=len(g)))]
        goff.append(len(umis))
    arr = sq.pack_batch([u.encode() for u in umis], klass=0)
    rep, ccounts, ncl = sq.umi_collapse(arr, np.array(counts, dtype=np.int64), np.array(goff, dtype=np.int64), threshold, method)
    rep, ccounts, ncl = rep.cpu().numpy(), ccounts.cpu().numpy(), ncl.cpu().numpy()
    for gi in range(len(group_sizes)):
        lo, hi = goff[gi], goff[gi + 1]
        want = umi_tools_groups(umis[lo:hi], counts[lo:hi], threshold, method)
        assert ncl[gi] == len(want)
        got = {}
        for v in range(lo, hi):
            got.setdefault(int(rep[v]), []).append(v - lo)
        assert sorted(sorted(g) for g in got.values()) == sorted(sorted(g) for g in want)
        for r, members in got.items():                    # the representative is a most frequent member, and carries the sum
            assert lo <= r < hi and counts[r] == max(counts[lo + m] for m in members)
            assert ccounts[r] == sum(counts[lo + m] for m in members)
    assert ccounts.sum() == sum(counts)


def test_umi_collapse_large_group_and_mixed_lengths(sq):
    rng = np.random.default_rng(99)
    n = 4400                                               # larger than the shared-memory staging: global-memory path
    seen = set()
    while len(seen) < n:
        seen.add("".join(rng.choice(list("ACGT"), size=int(rng.choice([8, 12])))))
    umis = sorted(seen)
    counts = [int(x) for x in rng.integers(1, 50, size=n)]
    arr = sq.pack_batch([u.encode() for u in umis], klass=0)
    rep, _, ncl = sq.umi_collapse(arr, np.array(counts), None, 1, "directional")
    want = umi_tools_groups(umis, counts, 1, "directional")
    got = {}
    for v, r in enumerate(rep.cpu().numpy().tolist()):
        got.setdefault(r, []).append(v)
    assert int(ncl[0]) == len(want)
    assert sorted(sorted(g) for g in got.values()) == sorted(sorted(g) for g in want)
