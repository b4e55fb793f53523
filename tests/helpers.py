"""Shared parity helpers: tie-aware top-k comparison and a numpy interpreter of the device
scan program (test-only; mirrors scan_kernel in deltapq_b200/csrc/kernels.cu)."""
import numpy as np

REL_TOL = 1e-5  # BASELINE.json north_star: ties within 1e-5 relative distance


def assert_topk_equal(pos_a, dist_a, pos_b, dist_b, node_dist=None, rel=REL_TOL):
    """Distances equal within `rel`; id sets equal after removing ids that tie (within `rel`)
    with the k-th distance; never compares id order inside a tie group."""
    pos_a, pos_b = np.asarray(pos_a, np.int64), np.asarray(pos_b, np.int64)
    dist_a, dist_b = np.asarray(dist_a, np.float64), np.asarray(dist_b, np.float64)
    assert pos_a.shape == pos_b.shape
    np.testing.assert_allclose(dist_a, dist_b, rtol=rel, atol=0)
    if pos_a.ndim == 1:
        pos_a, pos_b, dist_a, dist_b = pos_a[None], pos_b[None], dist_a[None], dist_b[None]
    for q in range(pos_a.shape[0]):
        kth = max(dist_a[q, -1], dist_b[q, -1])
        strict_a = {int(p) for p, d in zip(pos_a[q], dist_a[q]) if d < kth * (1 - rel)}
        strict_b = {int(p) for p, d in zip(pos_b[q], dist_b[q]) if d < kth * (1 - rel)}
        assert strict_a == strict_b, (q, sorted(strict_a ^ strict_b))
        if node_dist is not None:  # every reported id really has the reported distance
            nd = np.asarray(node_dist[q] if np.ndim(node_dist) == 2 else node_dist, np.float64)
            np.testing.assert_allclose(nd[pos_a[q]], dist_a[q], rtol=rel)


def interpret_program2(prog, table_q):
    """Code-array program (dpq_internal.h "v2", scan2.cu / scan8.cu) for ONE query on the CPU, the
    way the kernels read it: every node is `cstride` code bytes (pad bytes 0), the scan table has
    cstride * 256 rows, row = m * 256 + centroid, rows of subspaces >= M / centroids >= K all zero.
    Returns (positions, distances)."""
    M, K = prog["M"], prog["K"]
    nf = prog["v2_nf"]
    cp = prog["codes_padded"]
    assert prog["cstride"] == nf == cp.shape[1] and np.all(cp[:, M:] == 0)
    t = np.asarray(table_q)
    scan_table = np.zeros((nf, 256), t.dtype)
    scan_table[:M, :K] = t.reshape(M, K)
    d = scan_table[np.arange(nf)[None, :], cp].sum(1)
    pos = prog["base_pos"] + np.arange(prog["n_local"], dtype=np.int64)
    return pos, d


def interpret_any(prog, table_q):
    if prog.get("v2"):
        return interpret_program2(prog, table_q)
    return interpret_program(prog, table_q)


def interpret_program(prog, table_q):
    """Runs the record program for ONE query on the CPU.  table_q: integer or float table
    [M*K].  Returns (positions, distances) of every node the program emits, in order."""
    rb = prog["rb"]
    fmask = ((1 << rb) - 1) << 2
    tsh = rb + 2
    M = prog["codes"].shape[1]
    K = len(table_q) // M
    ops, chunks, anc = prog["ops"], prog["chunks"], prog["anc"]
    out_pos, out_d = [], []
    for c in range(chunks.shape[0]):
        qb, nq, first_pos, naf = (int(v) for v in chunks[c])
        n_anc, emit_root = naf & 0xFF, bool(naf & 0x100)
        stack = {}
        par = 0
        for lev in range(n_anc):
            d = sum(table_q[m * K + int(anc[c, lev, m])] for m in range(M))
            stack[lev] = d
            par = d
            if lev == 0 and emit_root:
                out_pos.append(0)
                out_d.append(d)
        pos = first_pos
        p, end = qb * 4, (qb + nq) * 4
        while p < end:
            w0 = int(ops[p])
            rq = (w0 >> 30) + 1
            d = par
            for w in ops[p:p + 4 * rq]:
                w = int(w)
                d = d + table_q[((w >> tsh) & fmask) >> 2] - table_q[(w & fmask) >> 2]
            out_pos.append(pos)
            out_d.append(d)
            lev = (w0 & 3) | ((w0 >> rb) & 0xC)
            if w0 & (1 << 29):
                par = d
                if w0 & (1 << 28):
                    stack[lev] = d
            elif w0 & (1 << 28):
                par = stack[lev]
            pos += 1
            p += 4 * rq
        assert p == end
    return np.array(out_pos, np.int64), np.array(out_d)


def gt_fixture(name):
    import hashlib
    import os
    import datagen as dg
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    gen = dg.gist_like if str(z["kind"]) == "gist" else dg.sift_like
    base = gen(int(z["n"]), int(z["d"]), seed=int(z["seed"]))
    assert hashlib.sha1(base.tobytes()).hexdigest() == str(z["base_sha1"]), "datagen drifted: regenerate the fixture"
    return base, z["queries"], int(z["topk"]), z["ref_ids"], z["ref_dist_text"]


def assert_gt_equals_reference_text(ids, dist, ref_ids, ref_text):
    """ids equal the reference's (modulo ties within the text precision); distances equal what the
    reference PRINTED: default ostream float formatting, 6 significant digits (pqbase.cpp:294-312)."""
    printed = np.array([[float("%g" % v) for v in row] for row in dist])
    np.testing.assert_allclose(printed, ref_text, rtol=2e-6)
    for q in range(len(ids)):
        if not np.array_equal(ids[q], ref_ids[q]):  # only neighbours whose printed distances tie may swap
            diff = ids[q] != ref_ids[q]
            assert set(ids[q][diff]) == set(ref_ids[q][diff]), q        # the same neighbours ...
            assert len(set(ref_text[q][diff])) < int(diff.sum()), q         # ... swapped inside a printed-distance tie
