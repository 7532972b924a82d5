"""ctypes binding of the C ABI declared in include/roargraph_b200.h (libroargraph_b200.so).

This is harness plumbing for tests/ and bench.py; the drop-in host layer is the C++ class
efanna2e::IndexBipartite in mysteryann_b200/host/.  There is no CPU fallback: if the CUDA library is
missing or no device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RG_B200_LIB") or os.path.join(PKG, "libroargraph_b200.so")  # RG_B200_LIB: A/B builds (tools/)

RG_OK = 0
RG_ERR_NOT_ENOUGH_RESULTS = 3
RG_ERR_NO_DEVICE = 5
METRIC_L2, METRIC_IP, METRIC_COSINE = 0, 1, 4

SYMBOLS = ["rg_last_error_string", "rg_version_string", "rg_device_count", "rg_index_create", "rg_index_destroy",
           "rg_index_info", "rg_search_batch", "rg_search_batch_device", "rg_search_expanded_device", "rg_search_configure", "rg_search_set_option", "rg_search_last_overflow_count", "rg_search_last_exception_count",
           "rg_host_register", "rg_host_unregister", "rg_index_launch_count", "rg_knn_exact", "rg_knn_exact_device", "rg_knn_merge_device", "rg_knn_merge",
           "rg_knn_last_stats", "rg_build_roargraph_device", "rg_build_roargraph", "rg_graph_info", "rg_graph_download", "rg_graph_destroy",
           "rg_index_create_from_graph", "rg_knn_last_second_pass_count", "rg_knn_release_scratch", "rg_knn_exact_sharded", "rg_knn_exact_grid", "rg_knn_exact_grid_host", "rg_knn_exact_sharded_host", "rg_build_projection_lists_device",
           "rg_knn_sharded_slice", "rg_nccl_get_unique_id", "rg_nccl_comm_init_rank", "rg_nccl_comm_init_all",
           "rg_nccl_comm_destroy", "rg_nccl_version"]

_lib = None


class RoarGraphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RoarGraphError(-1, f"{LIB_PATH} is missing: run `python -m mysteryann_b200.build` "
                                 "(the CUDA extension is required; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.rg_last_error_string.restype = C.c_char_p
    L.rg_version_string.restype = C.c_char_p
    L.rg_device_count.restype = i32
    L.rg_index_create.restype = i32
    L.rg_index_create.argtypes = [C.POINTER(vp), vp, u64, u32, i32, vp, vp, u32, i32, i32]
    L.rg_index_destroy.restype = i32
    L.rg_index_destroy.argtypes = [vp]
    L.rg_index_info.restype = i32
    L.rg_index_info.argtypes = [vp] + [vp] * 6
    L.rg_search_batch.restype = i32
    L.rg_search_batch.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp, vp]
    L.rg_search_batch_device.restype = i32
    L.rg_search_batch_device.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp, vp, vp, vp]
    L.rg_search_expanded_device.restype = i32
    L.rg_search_expanded_device.argtypes = [vp, u32, u64, u32, vp, vp, u32, vp]
    L.rg_search_configure.restype = i32
    L.rg_search_configure.argtypes = [vp, i32, i32, i32, i32, i32]
    L.rg_search_set_option.restype = i32
    L.rg_host_register.restype = i32
    L.rg_host_register.argtypes = [vp, C.c_uint64]
    L.rg_host_unregister.restype = i32
    L.rg_host_unregister.argtypes = [vp]
    L.rg_search_set_option.argtypes = [vp, C.c_char_p, i32]
    L.rg_search_last_overflow_count.restype = u32
    L.rg_search_last_overflow_count.argtypes = [vp]
    L.rg_search_last_exception_count.restype = u32
    L.rg_search_last_exception_count.argtypes = [vp]
    L.rg_index_launch_count.restype = u64
    L.rg_index_launch_count.argtypes = [vp]
    L.rg_knn_exact.restype = i32
    L.rg_knn_exact.argtypes = [vp, u64, u64, vp, u64, u32, i32, u32, vp, vp, i32]
    L.rg_knn_exact_device.restype = i32
    L.rg_knn_exact_device.argtypes = [vp, u64, u64, vp, u64, u32, i32, u32, vp, vp, i32, vp]
    L.rg_knn_merge_device.restype = i32
    L.rg_knn_merge_device.argtypes = [vp, vp, u32, u64, u32, i32, vp, vp, i32, vp]
    L.rg_knn_merge.restype = i32
    L.rg_knn_merge.argtypes = [vp, vp, u32, u64, u32, i32, vp, vp, i32]
    L.rg_knn_last_stats.restype = None
    L.rg_knn_last_stats.argtypes = [vp, vp]
    L.rg_knn_last_second_pass_count.restype = u64
    L.rg_knn_last_second_pass_count.argtypes = []
    L.rg_knn_release_scratch.restype = i32
    L.rg_knn_release_scratch.argtypes = [i32]
    L.rg_knn_exact_sharded.restype = i32
    L.rg_knn_exact_sharded.argtypes = [vp, u64, u64, vp, u64, u32, i32, u32, vp, vp, vp, i32, i32, i32, vp]
    L.rg_knn_exact_grid.restype = i32
    L.rg_knn_exact_grid.argtypes = [vp, u64, u64, vp, u64, u32, i32, u32, vp, vp, vp, i32, i32, i32, i32, vp]
    L.rg_knn_exact_sharded_host.restype = i32
    L.rg_knn_exact_sharded_host.argtypes = [vp, u64, u64, vp, u64, u32, i32, u32, vp, vp, vp, i32, i32, i32]
    L.rg_knn_sharded_slice.restype = None
    L.rg_knn_sharded_slice.argtypes = [u64, i32, i32, vp, vp]
    L.rg_nccl_get_unique_id.restype = i32
    L.rg_nccl_get_unique_id.argtypes = [vp]
    L.rg_nccl_comm_init_rank.restype = i32
    L.rg_nccl_comm_init_rank.argtypes = [C.POINTER(vp), i32, i32, vp, i32]
    L.rg_nccl_comm_init_all.restype = i32
    L.rg_nccl_comm_init_all.argtypes = [vp, i32, vp]
    L.rg_nccl_comm_destroy.restype = i32
    L.rg_nccl_comm_destroy.argtypes = [vp]
    L.rg_nccl_version.restype = i32
    L.rg_build_roargraph_device.restype = i32
    L.rg_build_roargraph_device.argtypes = [vp, u64, u32, i32, vp, u64, u32, u32, u32, u32, C.POINTER(vp), i32, vp]
    L.rg_build_projection_lists_device.restype = i32
    L.rg_build_projection_lists_device.argtypes = [vp, u64, u32, i32, vp, u64, u32, u32, u32, vp, i32, vp]
    L.rg_graph_info.restype = i32
    L.rg_graph_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rg_graph_download.restype = i32
    L.rg_graph_download.argtypes = [vp, vp, vp]
    L.rg_graph_destroy.restype = i32
    L.rg_graph_destroy.argtypes = [vp]
    L.rg_index_create_from_graph.restype = i32
    L.rg_index_create_from_graph.argtypes = [C.POINTER(vp), vp, u64, u32, i32, vp]
    _lib = L
    return L


def _check(rc):
    if rc != RG_OK:
        raise RoarGraphError(rc, (lib().rg_last_error_string() or b"").decode())


def device_count() -> int:
    return int(lib().rg_device_count())


def _hp(a):
    return None if a is None else a.ctypes.data


class Index:
    """Device-resident RoarGraph index (base rows + projection graph + entry point)."""

    @classmethod
    def from_graph(cls, d_base, graph, metric=METRIC_IP):
        """Search index over a device-built Graph; d_base is the CUDA torch tensor the graph was built from."""
        self = cls.__new__(cls)
        n, dim = d_base.shape
        h = C.c_void_p()
        _check(lib().rg_index_create_from_graph(C.byref(h), d_base.data_ptr(), n, dim, metric, graph._h))
        self._h, self._keep = h, d_base
        self.n, self.dim, self.metric, self.device, self.ep = int(n), int(dim), metric, d_base.device.index or 0, graph.ep
        return self

    def __init__(self, base, offsets, adj, ep, metric=METRIC_IP, device=0):
        """base: numpy float32 [n, dim] (copied to the device) or a CUDA torch tensor (adopted, kept alive)."""
        L = lib()
        offsets = np.ascontiguousarray(offsets, np.uint64)
        adj = np.ascontiguousarray(adj, np.uint32)
        self._keep = None
        if isinstance(base, np.ndarray):
            base = np.ascontiguousarray(base, np.float32)
            n, dim = base.shape
            ptr, on_dev = base.ctypes.data, 0
        else:  # torch CUDA tensor
            assert base.is_cuda and base.is_contiguous() and base.dtype.is_floating_point and base.element_size() == 4
            n, dim = base.shape
            ptr, on_dev = base.data_ptr(), 1
            device = base.device.index or 0
            self._keep = base
        self.n, self.dim, self.metric, self.device, self.ep = int(n), int(dim), metric, device, int(ep)
        h = C.c_void_p()
        _check(L.rg_index_create(C.byref(h), ptr, n, dim, metric, _hp(offsets), _hp(adj), ep, device, on_dev))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().rg_index_destroy(self._h)
            self._h = None

    __del__ = close

    def configure(self, gather=0, warps_per_query=0, ctas_per_sm=0, stage_rows=0, hash_log2=0, hash_space=0,
                  l2_hint=None, adj_prefetch=None, batch_mode=0):
        _check(lib().rg_search_configure(self._h, gather, warps_per_query, ctas_per_sm, stage_rows, hash_log2))
        _check(lib().rg_search_set_option(self._h, b"hash_space", hash_space))
        if l2_hint is not None:
            _check(lib().rg_search_set_option(self._h, b"l2_hint", l2_hint))
        if adj_prefetch is not None:
            _check(lib().rg_search_set_option(self._h, b"adj_prefetch", adj_prefetch))
        _check(lib().rg_search_set_option(self._h, b"batch_mode", batch_mode))

    def set_option(self, name: str, value: int):
        """Named option of rg_search_set_option ("hash_space", "l2_hint", "adj_prefetch", "batch_mode", "zero_copy")."""
        _check(lib().rg_search_set_option(self._h, name.encode(), int(value)))

    @property
    def last_overflow(self) -> int:
        return int(lib().rg_search_last_overflow_count(self._h))

    @property
    def last_exceptions(self) -> int:
        return int(lib().rg_search_last_exception_count(self._h))

    @property
    def launches(self) -> int:
        return int(lib().rg_index_launch_count(self._h))

    def search(self, queries, k, L, want_stats=True):
        """Host-buffer call (H2D + kernels + D2H inside).  queries: numpy float32 [nq, dim]."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq, d = queries.shape
        assert d == self.dim
        ids = np.empty((nq, k), np.uint32)
        dists = np.empty((nq, k), np.float32)
        cmps = np.empty(nq, np.uint32) if want_stats else None
        hops = np.empty(nq, np.uint32) if want_stats else None
        rc = lib().rg_search_batch(self._h, _hp(queries), nq, k, L, _hp(ids), _hp(dists), _hp(cmps), _hp(hops))
        if rc not in (RG_OK, RG_ERR_NOT_ENOUGH_RESULTS):
            _check(rc)
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, rc=rc)

    def search_raw(self, q_ptr, nq, k, L, ids_ptr, dists_ptr, cmps_ptr=None, hops_ptr=None):
        """Host-buffer call on raw addresses (e.g. pinned torch tensors)."""
        _check(lib().rg_search_batch(self._h, q_ptr, nq, k, L, ids_ptr, dists_ptr, cmps_ptr, hops_ptr))

    def search_device(self, d_queries, k, L, d_ids, d_dists, d_cmps=None, d_hops=None, d_status=None, stream=None):
        """Device-buffer call on CUDA torch tensors; asynchronous on `stream` (int handle or None)."""
        nq = d_queries.shape[0]
        p = lambda t: None if t is None else t.data_ptr()
        _check(lib().rg_search_batch_device(self._h, p(d_queries), nq, k, L, p(d_ids), p(d_dists), p(d_cmps),
                                            p(d_hops), p(d_status), stream))

    def search_expanded(self, node_lo, count, L, cap):
        """Build-time beam searches of base rows [node_lo, node_lo + count) (rg_search_expanded_device): returns
        (ids [count, cap] uint32, dists [count, cap] float32, counts [count]) of the expanded nodes in expansion order."""
        import torch

        keys = torch.zeros((count, cap), dtype=torch.int64, device=f"cuda:{self.device}")
        cnt = torch.zeros(count, dtype=torch.int32, device=f"cuda:{self.device}")
        _check(lib().rg_search_expanded_device(self._h, node_lo, count, L, keys.data_ptr(), cnt.data_ptr(), cap, None))
        torch.cuda.synchronize()
        k = keys.cpu().numpy().view(np.uint64)
        ids = ((k >> np.uint64(1)) & np.uint64(0x7FFFFFFF)).astype(np.uint32)
        o = (k >> np.uint64(32)).astype(np.uint32)                       # monotone image of the FP32 distance
        bits = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o)
        return ids, bits.astype(np.uint32).view(np.float32), cnt.cpu().numpy().view(np.uint32)


class Graph:
    """RoarGraph built on the GPU (rg_build_roargraph_device): fixed-stride adjacency + entry point, device resident."""

    def __init__(self, d_base, d_knn_ids, M_sq=100, M_pjbp=35, L_pjpq=500, metric=METRIC_IP, stream=None):
        """d_base: CUDA float32 [n, dim]; d_knn_ids: CUDA int32/uint32 [n_train, K] learn->base nearest neighbours."""
        n, dim = d_base.shape
        n_train, K = d_knn_ids.shape
        assert d_base.is_contiguous() and d_knn_ids.is_contiguous() and d_knn_ids.element_size() == 4
        h = C.c_void_p()
        _check(lib().rg_build_roargraph_device(d_base.data_ptr(), n, dim, metric, d_knn_ids.data_ptr(), n_train, K, M_sq,
                                               M_pjbp, L_pjpq, C.byref(h), d_base.device.index or 0, stream))
        self._h = h
        n_, md, nnz, ep = C.c_uint64(0), C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
        sec = (C.c_double * 6)()
        _check(lib().rg_graph_info(h, C.byref(n_), C.byref(md), C.byref(nnz), C.byref(ep), sec))
        self.n, self.max_degree, self.nnz, self.ep = n_.value, md.value, nnz.value, ep.value
        self.phase_seconds = dict(zip(("ep", "projection", "reverse", "enh_search", "enh_prune", "merge"), list(sec)))

    def download(self):
        """-> (ep, offsets u64 [n+1], adj u32 [nnz]) on the host (the CSR the index file format stores)."""
        off = np.empty(self.n + 1, np.uint64)
        adj = np.empty(self.nnz, np.uint32)
        _check(lib().rg_graph_download(self._h, _hp(off), _hp(adj)))
        return self.ep, off, adj

    def close(self):
        if getattr(self, "_h", None):
            lib().rg_graph_destroy(self._h)
            self._h = None

    __del__ = close


def projection_lists_device(d_base, d_knn_ids, M_sq=100, M_pjbp=35, metric=METRIC_IP):
    """P1 of the GPU build only: per-training-query pruned lists, torch int32 [n_train, M_pjbp + 1] (column 0 = length)."""
    import torch

    n, dim = d_base.shape
    n_train, K = d_knn_ids.shape
    out = torch.zeros((n_train, M_pjbp + 1), dtype=torch.int32, device=d_base.device)
    _check(lib().rg_build_projection_lists_device(d_base.data_ptr(), n, dim, metric, d_knn_ids.data_ptr(), n_train, K, M_sq, M_pjbp,
                                                  out.data_ptr(), d_base.device.index or 0, None))
    return out


def knn_exact(base, queries, K, metric=METRIC_IP, id_base=0, device=0):
    """Exact kNN of numpy host arrays (H2D + kernels + D2H inside).  Returns (ids u32 [nq,K], dists f32 [nq,K])."""
    base = np.ascontiguousarray(base, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    n, dim = base.shape
    nq = queries.shape[0]
    ids = np.empty((nq, K), np.uint32)
    dists = np.empty((nq, K), np.float32)
    _check(lib().rg_knn_exact(_hp(base), n, id_base, _hp(queries), nq, dim, metric, K, _hp(ids), _hp(dists), device))
    return ids, dists


def knn_exact_device(d_base, d_queries, K, d_ids, d_dists, metric=METRIC_IP, id_base=0, stream=None):
    """Exact kNN on CUDA torch tensors (float32 [n,dim], [nq,dim]; int32 [nq,K], float32 [nq,K])."""
    n, dim = d_base.shape
    device = d_base.device.index or 0
    _check(lib().rg_knn_exact_device(d_base.data_ptr(), n, id_base, d_queries.data_ptr(), d_queries.shape[0], dim,
                                     metric, K, d_ids.data_ptr(), d_dists.data_ptr(), device, stream))


def knn_merge_device(d_part_ids, d_part_dists, d_ids, d_dists, metric=METRIC_IP, stream=None):
    """K4: merge [G, nq, K] per-shard lists into the global top-K."""
    G, nq, K = d_part_ids.shape
    device = d_part_ids.device.index or 0
    _check(lib().rg_knn_merge_device(d_part_ids.data_ptr(), d_part_dists.data_ptr(), G, nq, K, metric,
                                     d_ids.data_ptr(), d_dists.data_ptr(), device, stream))


def knn_last_stats():
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib().rg_knn_last_stats(C.byref(a), C.byref(b))
    return dict(launches=a.value, exact_scans=b.value, second_pass=int(lib().rg_knn_last_second_pass_count()))


def knn_release_scratch(device=-1):
    _check(lib().rg_knn_release_scratch(device))


def knn_sharded_slice(nq, rank, world):
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    lib().rg_knn_sharded_slice(nq, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().rg_nccl_get_unique_id(buf))
    return buf.raw


def nccl_comm_init_rank(world, rank, unique_id: bytes, device):
    h = C.c_void_p()
    _check(lib().rg_nccl_comm_init_rank(C.byref(h), world, rank, C.c_char_p(unique_id), device))
    return h


def nccl_comm_destroy(comm):
    _check(lib().rg_nccl_comm_destroy(comm))


def knn_exact_grid(d_base_shard, id_base, d_group_queries, K, d_ids, d_dists, comm, rank, world, base_shards, metric=METRIC_IP,
                   stream=None):
    """rg_knn_exact_grid on CUDA torch tensors: rank holds base shard rank % base_shards and the queries of group
    rank // base_shards; d_ids/d_dists hold slice rank % base_shards of the group's queries."""
    n, dim = d_base_shard.shape
    device = d_base_shard.device.index or 0
    _check(lib().rg_knn_exact_grid(d_base_shard.data_ptr(), n, id_base, d_group_queries.data_ptr(), d_group_queries.shape[0], dim,
                                   metric, K, d_ids.data_ptr(), d_dists.data_ptr(), comm, rank, world, base_shards, device, stream))


def knn_exact_sharded(d_base_shard, id_base, d_queries, K, d_ids, d_dists, comm, rank, world, metric=METRIC_IP, stream=None):
    """rg_knn_exact_sharded on CUDA torch tensors; d_ids/d_dists hold this rank's query slice (knn_sharded_slice)."""
    n, dim = d_base_shard.shape
    device = d_base_shard.device.index or 0
    _check(lib().rg_knn_exact_sharded(d_base_shard.data_ptr(), n, id_base, d_queries.data_ptr(), d_queries.shape[0], dim,
                                      metric, K, d_ids.data_ptr(), d_dists.data_ptr(), comm, rank, world, device, stream))
