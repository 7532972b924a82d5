"""On-disk formats of the reference, byte for byte (numpy side; the C++ side is host/efanna2e/util.h).

* fbin  : <u32 n><u32 d><f32 n*d>                       (/root/reference include/efanna2e/util.h:106-127, 179-211)
* ibin  : <u32 n><u32 k><u32 ids n*k><f32 dists n*k>    (the "truthset" written by
          thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:325-343 and read by
          util.h:84-105,129-155 and src/index_bipartite.cpp:2622-2639 (ids only))
* index : <u32 ep><u32 n>{<u32 deg><u32 ids deg>} * n   (src/index_bipartite.cpp:2097-2117, 2606-2619)
"""
from __future__ import annotations

import numpy as np


def write_fbin(path, x: np.ndarray) -> None:
    x = np.ascontiguousarray(x, dtype=np.float32)
    with open(path, "wb") as f:
        np.array(x.shape, dtype=np.uint32).tofile(f)
        x.tofile(f)


def read_fbin(path) -> np.ndarray:
    with open(path, "rb") as f:
        n, d = np.fromfile(f, dtype=np.uint32, count=2)
        x = np.fromfile(f, dtype=np.float32, count=int(n) * int(d))
    if x.size != int(n) * int(d):
        raise RuntimeError("Data file size wrong!")  # util.h:124
    return x.reshape(int(n), int(d))


def write_ibin(path, ids: np.ndarray, dists: np.ndarray) -> None:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    dists = np.ascontiguousarray(dists, dtype=np.float32)
    assert ids.shape == dists.shape and ids.ndim == 2
    with open(path, "wb") as f:
        np.array(ids.shape, dtype=np.uint32).tofile(f)
        ids.tofile(f)
        dists.tofile(f)


def read_ibin(path):
    with open(path, "rb") as f:
        n, k = (int(v) for v in np.fromfile(f, dtype=np.uint32, count=2))
        ids = np.fromfile(f, dtype=np.uint32, count=n * k)
        dists = np.fromfile(f, dtype=np.float32, count=n * k)
    if ids.size != n * k or dists.size != n * k:
        raise RuntimeError("Data file size wrong!")  # util.h:102,152
    return ids.reshape(n, k), dists.reshape(n, k)


def write_index(path, ep: int, offsets: np.ndarray, adj: np.ndarray) -> None:
    """CSR (offsets u64[n+1], adj u32[]) -> projection index file."""
    n = len(offsets) - 1
    deg = np.diff(offsets).astype(np.uint32)
    out = np.empty(2 + n + len(adj), dtype=np.uint32)
    out[0], out[1] = ep, n
    # position of each degree word: 2 + i + offsets[i]
    pos = 2 + np.arange(n, dtype=np.int64) + offsets[:-1].astype(np.int64)
    out[pos] = deg
    mask = np.ones(out.size, dtype=bool)
    mask[:2] = False
    mask[pos] = False
    out[mask] = adj
    out.tofile(path)


def read_index(path):
    """projection index file -> (ep, offsets u64[n+1], adj u32[])."""
    return parse_index(np.fromfile(path, dtype=np.uint32))


def parse_index(raw: np.ndarray):
    ep, n = int(raw[0]), int(raw[1])
    body = raw[2:]
    # walk the degree words; vectorised by iterating in numpy-friendly chunks is awkward, so do a
    # tight python loop only over n (fine up to a few million) with a C fallback in the host lib.
    deg = np.empty(n, dtype=np.uint32)
    pos = 0
    for i in range(n):
        d = int(body[pos])
        deg[i] = d
        pos += 1 + d
    if pos != body.size:
        raise RuntimeError("index file size wrong")
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(deg, out=offsets[1:])
    degpos = np.arange(n, dtype=np.int64) + offsets[:-1].astype(np.int64)
    mask = np.ones(body.size, dtype=bool)
    mask[degpos] = False
    adj = np.ascontiguousarray(body[mask])
    return ep, offsets, adj


def pad_rows(x: np.ndarray, factor: int = 8) -> np.ndarray:
    """data_align (util.h:37-75): zero-pad the row length to a multiple of 8 floats."""
    n, d = x.shape
    nd = (d + factor - 1) // factor * factor
    if nd == d:
        return np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((n, nd), dtype=np.float32)
    out[:, :d] = x
    return out
