"""ctypes binding of libroargraph_host.so (the drop-in C++ host layer; mysteryann_b200/host/)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libroargraph_host.so")
BIN_DIR = os.path.join(PKG, "host", "bin")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.rgh_last_error.restype = C.c_char_p
        _lib.rgh_build_index.restype = C.c_int
        _lib.rgh_build_index.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int,
                                         C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_char_p, C.c_int, C.c_void_p]
        _lib.rgh_search_per_query.restype = C.c_int
        _lib.rgh_search_per_query.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                              C.c_uint32] + [C.c_void_p] * 5
        _lib.rgh_projection_lists.restype = C.c_int
        _lib.rgh_projection_lists.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                              C.c_uint32, C.c_void_p]
    return _lib


def projection_lists(base, knn_ids, metric=1, M_sq=100, M_pjbp=35):
    """P1 of the host build: the pruned pivot list of every training query, uint32 [n_train, M_pjbp + 1] (column 0 = length)."""
    base = np.ascontiguousarray(base, np.float32)
    knn_ids = np.ascontiguousarray(knn_ids, np.uint32)
    out = np.zeros((knn_ids.shape[0], M_pjbp + 1), np.uint32)
    rc = lib().rgh_projection_lists(base.ctypes.data, base.shape[1], metric, knn_ids.ctypes.data, knn_ids.shape[0],
                                    knn_ids.shape[1], M_sq, M_pjbp, out.ctypes.data)
    if rc:
        raise RuntimeError(lib().rgh_last_error().decode())
    return out


def build_index(base, train, knn_ids, out_path, metric=1, M_sq=100, M_pjbp=35, L_pjpq=500, threads=1, quiet=True):
    """BuildRoarGraph on the host CPU (same rules as the reference); writes the projection index file."""
    base = np.ascontiguousarray(base, np.float32)
    train = np.ascontiguousarray(train, np.float32)
    knn_ids = np.ascontiguousarray(knn_ids, np.uint32)
    assert base.shape[1] % 8 == 0 and base.shape[1] == train.shape[1] and knn_ids.shape[0] == train.shape[0]
    sec = C.c_double(0)
    rc = lib().rgh_build_index(base.ctypes.data, base.shape[0], train.ctypes.data, train.shape[0], base.shape[1],
                               metric, knn_ids.ctypes.data, knn_ids.shape[1], M_sq, M_pjbp, L_pjpq, threads,
                               out_path.encode(), int(quiet), C.byref(sec))
    if rc:
        raise RuntimeError(lib().rgh_last_error().decode())
    return sec.value


def search_per_query(base_fbin, index_file, queries, k, L_pq, metric=1, threads=8):
    """The reference driver's loop on the drop-in class: one IndexBipartite::SearchRoarGraph call per query from `threads`
    OpenMP threads (concurrent callers are micro-batched into one GPU launch inside the class).  Returns dict(ids, dists, cmps,
    hops, seconds = wall time of the loop)."""
    queries = np.ascontiguousarray(queries, np.float32)
    nq = queries.shape[0]
    assert queries.shape[1] % 8 == 0
    ids = np.empty((nq, k), np.uint32)
    dists = np.empty((nq, k), np.float32)
    cmps = np.empty(nq, np.uint32)
    hops = np.empty(nq, np.uint32)
    sec = C.c_double(0)
    rc = lib().rgh_search_per_query(str(base_fbin).encode(), str(index_file).encode(), metric, queries.ctypes.data, nq, k,
                                    L_pq, threads, ids.ctypes.data, dists.ctypes.data, cmps.ctypes.data, hops.ctypes.data,
                                    C.byref(sec))
    if rc:
        raise RuntimeError(lib().rgh_last_error().decode())
    return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, seconds=sec.value)
