"""Deterministic synthetic SIFT/GIST-shaped data + file writers (test / bench tooling).

Formats follow SURVEY.md App. A (reference: utils.cpp:14-71 fvecs/bvecs, pq.cpp:267-286
codebook text, pq_tree.cpp:1011-1031 codes file).  Nothing here is on the product path:
the same files feed the reference binaries, the oracle and the CUDA library.
"""
import os
import numpy as np


def sift_like(n, d=128, seed=0, n_clusters=256, sigma=14.0, chunk=200_000):
    """Non-negative integer-valued float32 vectors, 0..255, clustered per 16-d block.

    Each 16-d block draws its own cluster id (product structure, like real SIFT after PQ),
    with a Zipf-ish popularity so PQ codes share many sub-codes but few are exact
    duplicates.  Calibrated so the 1M / M=8 / K=256 DeltaTree has 3.5-5 changed subspaces
    per node.
    """
    rng = np.random.default_rng(seed)
    nb = max(1, d // 16)
    crng = np.random.default_rng(1234567)  # centres are shared by base / learn / query
    centres = np.clip(crng.gamma(2.0, 22.0, size=(nb, n_clusters, (d + nb - 1) // nb)), 0, 255)
    pop = 1.0 / np.arange(1, n_clusters + 1) ** 0.7
    pop /= pop.sum()
    out = np.empty((n, d), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        for b in range(nb):
            lo, hi = b * (d // nb), (b + 1) * (d // nb) if b < nb - 1 else d
            cid = rng.choice(n_clusters, size=e - s, p=pop)
            x = centres[b, cid, : hi - lo] + rng.normal(0.0, sigma, size=(e - s, hi - lo))
            out[s:e, lo:hi] = np.clip(np.rint(x), 0, 255)
    return out


def gist_like(n, d=960, seed=0, n_clusters=128, sigma=0.04, chunk=50_000):
    """float32 in [0,1), not integer valued (exercises FP rounding in encode / groundtruth)."""
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(7654321)
    nb = 16
    w = d // nb
    centres = crng.random((nb, n_clusters, w)) * 0.5
    out = np.empty((n, d), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        for b in range(nb):
            cid = rng.integers(0, n_clusters, size=e - s)
            x = centres[b, cid] + rng.normal(0.0, sigma, size=(e - s, w))
            out[s:e, b * w:(b + 1) * w] = np.clip(x, 0, 0.999).astype(np.float32)
    return out


def kmeans_codebook(learn, M, K, iters=8, seed=0):
    """Per-subspace Lloyd k-means -> codebook [M][K][Ds] float32 (an INPUT to every parity
    checked stage; the reference's own learn task is RNG-driven)."""
    n, D = learn.shape
    Ds = (D + M - 1) // M
    if Ds * M != D:
        learn = np.concatenate([learn, np.zeros((n, Ds * M - D), np.float32)], axis=1)
    rng = np.random.default_rng(seed)
    cw = np.empty((M, K, Ds), dtype=np.float32)
    for m in range(M):
        sub = learn[:, m * Ds:(m + 1) * Ds].astype(np.float64)
        uniq = np.unique(sub, axis=0)
        src = uniq if len(uniq) >= K else sub
        c = src[rng.choice(len(src), size=K, replace=len(src) < K)].copy()
        for _ in range(iters):
            d2 = (sub * sub).sum(1)[:, None] - 2.0 * sub @ c.T + (c * c).sum(1)[None, :]
            a = d2.argmin(1)
            for k in range(K):
                sel = sub[a == k]
                if len(sel):
                    c[k] = sel.mean(0)
                else:
                    c[k] = sub[rng.integers(len(sub))] + rng.normal(0, 1e-3, size=Ds)
        cw[m] = c.astype(np.float32)
    return cw


def write_vecs(path, x, ext="fvecs"):
    n, d = x.shape
    if ext == "fvecs":
        rec = np.empty((n, d + 1), dtype=np.float32)
        rec[:, 0] = np.array([d], dtype=np.int32).view(np.float32)[0]
        rec[:, 1:] = x
        rec.tofile(path)
    elif ext == "bvecs":
        rec = np.empty((n, d + 4), dtype=np.uint8)
        rec[:, :4] = np.frombuffer(np.int32(d).tobytes(), dtype=np.uint8)
        rec[:, 4:] = x.astype(np.uint8)
        rec.tofile(path)
    else:
        raise ValueError(ext)


def read_vecs(path, ext="fvecs"):
    if ext == "fvecs":
        a = np.fromfile(path, dtype=np.float32)
        d = int(a[:1].view(np.int32)[0])
        return a.reshape(-1, d + 1)[:, 1:].copy()
    a = np.fromfile(path, dtype=np.uint8)
    d = int(a[:4].view(np.int32)[0])
    return a.reshape(-1, d + 4)[:, 4:].astype(np.float32)


def write_codebook(path, cw):
    """pq.cpp:267-286: 'M,Ks,Ds' then per m 'm:' and Ks lines of Ds '%g,' values."""
    M, K, Ds = cw.shape
    with open(path, "w") as f:
        f.write(f"{M},{K},{Ds}\n")
        for m in range(M):
            f.write(f"{m}:\n")
            for k in range(K):
                f.write("".join("%g," % v for v in cw[m, k]) + "\n")


def read_codebook(path):
    """pq.cpp:288-312 (values are what survives the 6-digit text round trip)."""
    with open(path) as f:
        M, K, Ds = (int(v) for v in f.readline().strip().split(","))
        cw = np.empty((M, K, Ds), dtype=np.float32)
        for m in range(M):
            assert f.readline().strip() == f"{m}:"
            for k in range(K):
                vals = f.readline().strip().rstrip(",").split(",")
                cw[m, k] = np.array([float(v) for v in vals], dtype=np.float32)
    return cw


def roundtrip_codebook(cw):
    """Codebook as every stage sees it: after the %g text round trip."""
    flat = np.array([float("%g" % v) for v in cw.ravel()], dtype=np.float32)
    return flat.reshape(cw.shape)


def write_codes(path, codes):
    """pq_tree.cpp:1011-1031: int64 N then N*M bytes."""
    with open(path, "wb") as f:
        f.write(np.int64(codes.shape[0]).tobytes())
        f.write(np.ascontiguousarray(codes, dtype=np.uint8).tobytes())


def read_codes(path, M):
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
        return np.frombuffer(f.read(n * M), dtype=np.uint8).reshape(n, M).copy()


def read_dtc(path):
    """Compressed DeltaTree file -> (n_codes, payload bytes)."""
    raw = np.fromfile(path, dtype=np.uint8)
    n_codes, n_bytes = (int(v) for v in raw[:16].view(np.int64))
    return n_codes, raw[16:16 + n_bytes].copy()


def make_dataset(dirpath, n, n_query, M=8, K=256, d=128, kind="sift", seed=0, n_learn=None):
    """Writes base/query fvecs + codebook into dirpath; returns (base, queries, cw)."""
    os.makedirs(dirpath, exist_ok=True)
    gen = sift_like if kind == "sift" else gist_like
    base = gen(n, d, seed=seed + 1)
    queries = gen(n_query, d, seed=seed + 2)
    learn = gen(n_learn or min(max(4 * K, n), 20000), d, seed=seed + 3)
    cw = roundtrip_codebook(kmeans_codebook(learn, M, K, seed=seed))
    write_vecs(os.path.join(dirpath, "base.fvecs"), base)
    write_vecs(os.path.join(dirpath, "query.fvecs"), queries)
    write_codebook(os.path.join(dirpath, f"M{M}K{K}codewords.txt"), cw)
    return base, queries, cw
