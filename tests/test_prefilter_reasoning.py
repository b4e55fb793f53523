"""CPU: the invariant behind the edge search's futile-pass prefilter (edges.cu).

A pass with dropped-subspace set D merges something only if two pool codes agree on every kept
subspace.  Simulating the pass order of find_edges_by_diff_approx (DCAT.h:1207-1332: rounds by
number of dropped subspaces, every run keeps one survivor) shows that whenever a pass merges, D is
exactly the changed-subspace mask of some pair of ORIGINAL codes -- pairs whose mask is a strict
subset of D were separated in the earlier pass that dropped exactly their mask.  The GPU tests
compare the filtered edge search with the oracle and with the unfiltered run."""
import itertools

import numpy as np
import pytest


@pytest.mark.parametrize("seed,M,K,n,dup", [(3, 7, 6, 400, 0.2), (5, 8, 4, 300, 0.0), (9, 6, 9, 500, 0.05)])
def test_a_pass_merges_only_if_its_dropped_set_is_some_pairs_exact_mask(seed, M, K, n, dup):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, K, size=(n, M))
    codes[rng.random(n) < dup] = codes[0]
    weights = 1 << np.arange(M)
    masks = set()
    for i in range(n - 1):
        masks.update(((codes[i] != codes[i + 1:]) * weights).sum(axis=1).tolist())
    pool = list(range(n))
    merging_passes = 0
    for d in range(M + 1):
        sel = [1] * (M - d) + [0] * d
        for s in sorted(set(itertools.permutations(sel)), reverse=True):  # std::prev_permutation order
            kept = [m for m in range(M) if s[m]]
            dropped = sum(1 << m for m in range(M) if not s[m])
            runs = {}
            for v in pool:
                runs.setdefault(tuple(codes[v][kept]), []).append(v)
            if any(len(r) > 1 for r in runs.values()):
                merging_passes += 1
                assert dropped in masks
            # one survivor per run; WHICH member (the parent rule: largest height / first member)
            # does not matter for the invariant, so any choice must pass
            pool = sorted(r[int(rng.integers(0, len(r)))] for r in runs.values())
    assert len(pool) == 1 and merging_passes > 0
