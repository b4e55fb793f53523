"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for oracle/libdpq_oracle.so (the plain-C
restatement) and oracle/_ref/libref_harness.so (the UNMODIFIED reference behind a C shim).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")


def build(ref=True):
    """make oracle (+ ref when /root/reference is present)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libdpq_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.dpqo_lut.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.dpqo_scan.restype = C.c_int64
        L.dpqo_scan.argtypes = [_u8p, C.c_int64, C.c_int64, C.c_int, C.c_int, _f32p, C.c_int,
                                _i32p, _f32p, C.c_void_p]
        L.dpqo_decode.restype = C.c_int64
        L.dpqo_decode.argtypes = [_u8p, C.c_int64, C.c_int64, C.c_int, _u8p, _u8p, _i32p]
        L.dpqo_encode.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int64, C.c_int, _u8p]
        L.dpqo_find_edges.restype = C.c_int64
        L.dpqo_find_edges.argtypes = [_u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _u32p,
                                      C.POINTER(C.c_uint32)]
        L.dpqo_centroid_tables.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.dpqo_layout.argtypes = [_u8p, C.c_int64, C.c_int, C.c_int, _u32p, C.c_uint32, _f32p,
                                  _u32p, _u32p, _u32p, _u8p, _f32p, _f32p]
        L.dpqo_qnodes8.argtypes = [_u8p, C.c_int64, _u32p, _u32p, _u32p, _u8p, _f32p, _f32p, _u8p]
        L.dpqo_stream_bytes.restype = C.c_int64
        L.dpqo_stream_bytes.argtypes = [_u8p, C.c_int64, C.c_int, _u32p, _u32p]
        L.dpqo_stream.restype = C.c_int64
        L.dpqo_stream.argtypes = [_u8p, C.c_int64, C.c_int, _u32p, _u32p, _u8p, _u8p]
        L.dpqo_groundtruth_chunk.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, C.c_int, C.c_int,
                                             C.c_int, _f32p, _u32p, _i32p]
        L.dpqo_groundtruth_finish.argtypes = [C.c_int, C.c_int, _f32p, _u32p, _i32p, _u32p, _f32p]
        _lib = L
    return _lib


def lut(cw, query):
    M, K, Ds = cw.shape
    out = np.empty((M, K), np.float32)
    lib().dpqo_lut(np.ascontiguousarray(cw), M, K, Ds, np.ascontiguousarray(query, np.float32), out)
    return out


def scan(payload, n_codes, cw, query, topk, want_node_dist=False, lut_in=None):
    M, K, Ds = cw.shape
    table = lut(cw, query) if lut_in is None else np.ascontiguousarray(lut_in, np.float32)
    pos = np.empty(topk, np.int32)
    dist = np.empty(topk, np.float32)
    nd = np.empty(n_codes, np.float32) if want_node_dist else None
    used = lib().dpqo_scan(payload, len(payload), n_codes, M, K, table, topk, pos, dist,
                           nd.ctypes.data if nd is not None else None)
    assert used == len(payload), (used, len(payload))
    return (pos, dist, nd) if want_node_dist else (pos, dist)


def decode(payload, n_codes, M):
    codes = np.empty((n_codes, M), np.uint8)
    depth = np.empty(n_codes, np.uint8)
    parent = np.empty(n_codes, np.int32)
    used = lib().dpqo_decode(payload, len(payload), n_codes, M, codes, depth, parent)
    assert used == len(payload), (used, len(payload))
    return codes, depth, parent


def encode(cw, x):
    M, K, Ds = cw.shape
    x = np.ascontiguousarray(x, np.float32)
    codes = np.empty((x.shape[0], M), np.uint8)
    lib().dpqo_encode(np.ascontiguousarray(cw), M, K, Ds, x, x.shape[0], x.shape[1], codes)
    return codes


def find_edges(codes, K=256, h=1, method=1):
    n, M = codes.shape
    edges = np.zeros((max(n - 1, 1), 2), np.uint32)
    root = C.c_uint32(0)
    ne = lib().dpqo_find_edges(np.ascontiguousarray(codes), n, M, K, h, method, edges, C.byref(root))
    assert ne == n - 1, (ne, n)
    return edges[: n - 1], int(root.value)


def centroid_tables(cw):
    M, K, Ds = cw.shape
    t = np.empty((M, K, K), np.float32)
    lib().dpqo_centroid_tables(np.ascontiguousarray(cw), M, K, Ds, t)
    return t


def layout(codes, edges, root, tables, K=256):
    n, M = codes.shape
    out = dict(vec_id=np.empty(n, np.uint32), parent_pos=np.empty(n, np.uint32),
               child_num=np.empty(n, np.uint32), depth=np.empty(n, np.uint8),
               max_dist=np.empty(n, np.float32), max_dist2p=np.empty(n, np.float32))
    lib().dpqo_layout(np.ascontiguousarray(codes), n, M, K, np.ascontiguousarray(edges), root,
                      np.ascontiguousarray(tables), out["vec_id"], out["parent_pos"],
                      out["child_num"], out["depth"], out["max_dist"], out["max_dist2p"])
    return out


def qnodes8(codes, lay):
    n = codes.shape[0]
    buf = np.zeros((n + 1) * 60, np.uint8)
    lib().dpqo_qnodes8(np.ascontiguousarray(codes), n, lay["vec_id"], lay["parent_pos"],
                       lay["child_num"], lay["depth"], lay["max_dist"], lay["max_dist2p"], buf)
    return buf


def stream(codes, lay):
    n, M = codes.shape
    codes = np.ascontiguousarray(codes)
    nb = lib().dpqo_stream_bytes(codes, n, M, lay["vec_id"], lay["parent_pos"])
    payload = np.zeros(nb, np.uint8)
    used = lib().dpqo_stream(codes, n, M, lay["vec_id"], lay["parent_pos"], lay["depth"], payload)
    assert used == nb, (used, nb)
    return payload


def build_tree(codes, cw, h=1, method=1):
    """codes -> (edges, root, layout dict, payload): the whole approx_tree task."""
    K = cw.shape[1]
    edges, root = find_edges(codes, K, h, method)
    lay = layout(codes, edges, root, centroid_tables(cw), K)
    return edges, root, lay, stream(codes, lay)


def groundtruth(base, queries, topk, chunk=100000):
    base = np.ascontiguousarray(base, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    Q, D = queries.shape
    hd = np.zeros((Q, topk), np.float32)
    hi = np.zeros((Q, topk), np.uint32)
    cnt = np.zeros(Q, np.int32)
    for s in range(0, base.shape[0], chunk):
        blk = base[s:s + chunk]
        lib().dpqo_groundtruth_chunk(blk, blk.shape[0], s, queries, Q, D, topk, hd, hi, cnt)
    oid = np.empty((Q, topk), np.uint32)
    od = np.empty((Q, topk), np.float32)
    lib().dpqo_groundtruth_finish(Q, topk, hd, hi, cnt, oid, od)
    return oid, od


# ---------------------------------------------------------------- the real reference ---
REF_DIR = os.path.join(HERE, "_ref")
_ref = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_harness.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_harness.so"))
        L.ref_scan_in_memory.restype = C.c_double
        L.ref_scan_in_memory.argtypes = [_u8p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                         _f32p, _f32p, C.c_int, C.c_int, _i32p, _f32p, C.c_void_p]
        L.ref_encode.restype = C.c_double
        L.ref_encode.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_longlong, C.c_int, _u8p]
        _ref = L
    return _ref


def ref_scan(payload, n_codes, cw, queries, topk, want_lut=False):
    """Reference in-memory scan (DCAT.h:3731) for a batch -> pos[Q][k], dist[Q][k], seconds."""
    M, K, Ds = cw.shape
    queries = np.ascontiguousarray(queries, np.float32)
    Q = queries.shape[0]
    pos = np.empty((Q, topk), np.int32)
    dist = np.empty((Q, topk), np.float32)
    lut_out = np.empty((Q, M, K), np.float32) if want_lut else None
    secs = ref().ref_scan_in_memory(payload, len(payload), n_codes, M, K, Ds,
                                    np.ascontiguousarray(cw), queries, Q, topk, pos, dist,
                                    lut_out.ctypes.data if want_lut else None)
    return (pos, dist, secs, lut_out) if want_lut else (pos, dist, secs)


def ref_encode(cw, x):
    M, K, Ds = cw.shape
    x = np.ascontiguousarray(x, np.float32)
    codes = np.empty((x.shape[0], M), np.uint8)
    secs = ref().ref_encode(np.ascontiguousarray(cw), M, K, Ds, x, x.shape[0], x.shape[1], codes)
    return codes, secs
