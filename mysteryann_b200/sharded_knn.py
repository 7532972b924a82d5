"""Multi-GPU exact kNN: base rows sharded over the ranks, one exchange step, K4 merge.

Same decomposition as the reference's tool, which walks the base in parts of 20M points sequentially and merges the
per-part top-k (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:32, 396-448); here the parts are the ranks of
one 8xB200 box (one process per GPU).  Rank g holds base rows [lo_g, hi_g) and all queries and ends up with the merged
lists of query slice g.

Product path `knn_sharded`: one call of the C ABI entry rg_knn_exact_sharded (csrc/rg_knn_sharded.cu) per rank - K2/K3
per shard and query chunk, grouped ncclSend/ncclRecv of the per-shard lists over NVLink, K4 merge, all issued by the
library on its own NCCL communicator.  torch.distributed is only the side channel that hands the ncclUniqueId to the
ranks (`nccl_comm_for`) and, optionally, all-gathers the merged slices (the learn->base file is written by one rank).

Exchange volume per rank: (G-1)/G * nq * K * 8 bytes in and out (C4, nq = 10M, K = 100, G = 8: 7 GB, ~10 ms at the
measured 770 GB/s per direction) against ~5 PFLOP of GEMM per rank.

`knn_sharded_with` is the same algorithm with the two compute steps injected and torch.distributed's all-to-all as the
exchange: it runs on whatever device the tensors live on, so the shard bounds, the exchange layout and the gather are
covered by world_size-2/4 gloo tests on CPU (tests/test_sharded_knn_cpu.py, with the oracle standing in for the CUDA
kernels); `knn_sharded_torch` is that variant with the CUDA kernels (round 1's path, kept as a yardstick).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[int]:
    """Contiguous row ranges: shard g = [b[g], b[g+1]); the first n % world shards get one extra row."""
    if world <= 0:
        raise ValueError("world must be positive")
    q, r = divmod(int(n), world)
    b = [0]
    for g in range(world):
        b.append(b[-1] + q + (1 if g < r else 0))
    return b


def exchange_partials(part_ids: torch.Tensor, part_dists: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """All-to-all of per-shard lists.  part_*: [nq, K] on this rank (lists of ALL queries against this rank's shard).
    Returns ([G, nq_g, K] ids, [G, nq_g, K] dists, query bounds) where slice g = queries [qb[g], qb[g+1]) is the one
    this rank merges and index 0 of the result runs over the source ranks (= base shards, ascending ids)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nq, K = part_ids.shape
    qb = shard_bounds(nq, world)
    rows_in = [qb[g + 1] - qb[g] for g in range(world)]
    mine = rows_in[rank]
    rows_out = [mine] * world
    out_ids = torch.empty((world * mine, K), dtype=part_ids.dtype, device=part_ids.device)
    out_d = torch.empty((world * mine, K), dtype=part_dists.dtype, device=part_dists.device)
    dist.all_to_all_single(out_ids, part_ids.contiguous(), rows_out, rows_in, group=group)
    dist.all_to_all_single(out_d, part_dists.contiguous(), rows_out, rows_in, group=group)
    return out_ids.view(world, mine, K), out_d.view(world, mine, K), qb


def gather_rows(local: torch.Tensor, bounds: List[int], group=None) -> torch.Tensor:
    """All-gather of row slices of unequal length (slice g = rows [bounds[g], bounds[g+1])) into the full array."""
    world = dist.get_world_size(group)
    longest = max(bounds[g + 1] - bounds[g] for g in range(world))
    padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view((world, longest) + tuple(local.shape[1:]))
    return torch.cat([out[g, : bounds[g + 1] - bounds[g]] for g in range(world)], dim=0)


class Grid:
    """2-D decomposition of `world` ranks into `base_shards` x `query_groups`: rank r holds base shard r % base_shards and
    answers query group r // base_shards; the exchange + merge of `knn_sharded` then runs inside each query group
    (a process group of `base_shards` ranks).  base_shards == world is the plain base-sharded scheme of BASELINE.json's
    config C4; base_shards == 1 shards the queries only (no exchange at all) - the better choice when a base shard would
    be so small that the per-batch threshold warm-up of K2 dominates (1.25M-row shards reach 663
    TFLOP/s per GPU, 5M-row shards 914)."""

    def __init__(self, base_shards: int, world: Optional[int] = None, rank: Optional[int] = None):
        world = dist.get_world_size() if world is None else world
        rank = dist.get_rank() if rank is None else rank
        if base_shards <= 0 or world % base_shards:
            raise ValueError(f"base_shards={base_shards} must divide the world size {world}")
        self.world, self.rank = world, rank
        self.base_shards, self.query_groups = base_shards, world // base_shards
        self.shard, self.qgroup = rank % base_shards, rank // base_shards
        self.group = None

    def make_groups(self):
        """Collective: every rank creates every query group's process group (torch.distributed requires that) and keeps
        its own.  Returns self."""
        for g in range(self.query_groups):
            ranks = list(range(g * self.base_shards, (g + 1) * self.base_shards))
            pg = dist.new_group(ranks)
            if g == self.qgroup:
                self.group = pg
        return self

    def base_bounds(self, n: int) -> Tuple[int, int]:
        b = shard_bounds(n, self.base_shards)
        return b[self.shard], b[self.shard + 1]

    def query_bounds(self, nq: int) -> Tuple[int, int]:
        b = shard_bounds(nq, self.query_groups)
        return b[self.qgroup], b[self.qgroup + 1]

    def result_bounds(self, nq: int) -> List[int]:
        """Row bounds, in rank order, of the merged slices the ranks end up with (gather=False): query group by query
        group, inside a group slice by slice - contiguous and ascending, so `gather_rows(slice, bounds)` over the whole
        world assembles the full answer."""
        out = [0]
        qb = shard_bounds(nq, self.query_groups)
        for g in range(self.query_groups):
            inner = shard_bounds(qb[g + 1] - qb[g], self.base_shards)
            out += [qb[g] + v for v in inner[1:]]
        return out


def knn_sharded_with(local_knn: Callable, merge: Callable, queries: torch.Tensor, K: int, group=None,
                     gather: bool = True):
    """The sharded algorithm with the two compute steps injected (CUDA kernels in the product path, the oracle in the
    CPU tests): local_knn(queries) -> (ids [nq,K] global, dists [nq,K]); merge(ids [G,nq_g,K], dists) -> (ids, dists)."""
    ids, d = local_knn(queries)
    pid, pd, qb = exchange_partials(ids, d, group)
    mid, md = merge(pid, pd)
    if not gather:
        return mid, md, qb
    return gather_rows(mid, qb, group), gather_rows(md, qb, group), qb


_COMMS = {}  # process-group id -> (ncclComm_t handle, rank in group, group size)


def nccl_comm_for(group=None, device: Optional[int] = None):
    """The library's own NCCL communicator for `group` (created once): the 128-byte ncclUniqueId is made by the group's
    first rank through the C ABI (rg_nccl_get_unique_id) and handed to the others with a torch.distributed broadcast -
    torch only plays the side channel; the exchange itself is issued by rg_knn_exact_sharded.  Collective over the group."""
    from . import capi

    key = id(group) if group is not None else 0
    if key in _COMMS:
        return _COMMS[key]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = torch.cuda.current_device() if device is None else device
    src = dist.get_global_rank(group, 0) if group is not None else 0
    if dist.get_rank() == src:
        uid = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).clone()
    else:
        uid = torch.zeros(128, dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        uid = uid.cuda(device)
    dist.broadcast(uid, src=src, group=group)
    comm = capi.nccl_comm_init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()), device)
    _COMMS[key] = (comm, rank, world)
    return _COMMS[key]


def knn_sharded(d_base_shard: torch.Tensor, id_base: int, d_queries: torch.Tensor, K: int, metric: int = 1, group=None,
                gather: bool = True, stream: Optional[int] = None):
    """Product path (CUDA tensors): exact top-K of every query over the union of all ranks' base shards through the C ABI
    entry rg_knn_exact_sharded (K2/K3 per shard, grouped ncclSend/ncclRecv exchange, K4 merge - all inside the library).
    Returns (ids int32 [nq or nq_g, K], dists float32, query bounds).  Raises if the CUDA library or a device is missing."""
    from . import capi

    if not (d_base_shard.is_cuda and d_queries.is_cuda):
        raise capi.RoarGraphError(capi.RG_ERR_NO_DEVICE, "knn_sharded needs CUDA tensors (there is no CPU fallback)")
    comm, rank, world = nccl_comm_for(group, d_base_shard.device.index or 0)
    nq = d_queries.shape[0]
    qb = shard_bounds(nq, world)
    assert (qb[rank], qb[rank + 1]) == capi.knn_sharded_slice(nq, rank, world)
    mine = qb[rank + 1] - qb[rank]
    ids = torch.empty((mine, K), dtype=torch.int32, device=d_queries.device)
    d = torch.empty((mine, K), dtype=torch.float32, device=d_queries.device)
    capi.knn_exact_sharded(d_base_shard, id_base, d_queries, K, ids, d, comm, rank, world, metric=metric, stream=stream)
    if not gather:
        return ids, d, qb
    return gather_rows(ids, qb, group), gather_rows(d, qb, group), qb


def grid_layout(rank: int, world: int, base_shards: int, n: int, nq: int):
    """Rows of the base and of the query set rank `rank` works on in the base_shards x (world / base_shards) grid of
    rg_knn_exact_grid, and the query rows whose merged lists it ends up with:
    ((base_lo, base_hi), (group_lo, group_hi), (out_lo, out_hi)) - all as global row ranges."""
    assert base_shards >= 1 and world % base_shards == 0
    groups = world // base_shards
    s, g = rank % base_shards, rank // base_shards
    bb, gb = shard_bounds(n, base_shards), shard_bounds(nq, groups)
    sb = shard_bounds(gb[g + 1] - gb[g], base_shards)
    return (bb[s], bb[s + 1]), (gb[g], gb[g + 1]), (gb[g] + sb[s], gb[g] + sb[s + 1])


def knn_grid(d_base_shard: torch.Tensor, id_base: int, d_group_queries: torch.Tensor, K: int, base_shards: int, metric: int = 1,
             group=None, stream: Optional[int] = None):
    """rg_knn_exact_grid: the ranks form world / base_shards query groups of base_shards ranks; this rank holds base shard
    rank % base_shards and its group's queries (see grid_layout) and gets (ids, dists) of its slice of the group's queries."""
    from . import capi

    if not (d_base_shard.is_cuda and d_group_queries.is_cuda):
        raise capi.RoarGraphError(capi.RG_ERR_NO_DEVICE, "knn_grid needs CUDA tensors (there is no CPU fallback)")
    comm, rank, world = nccl_comm_for(group, d_base_shard.device.index or 0)
    sb = shard_bounds(d_group_queries.shape[0], base_shards)
    mine = sb[rank % base_shards + 1] - sb[rank % base_shards]
    ids = torch.empty((mine, K), dtype=torch.int32, device=d_group_queries.device)
    d = torch.empty((mine, K), dtype=torch.float32, device=d_group_queries.device)
    capi.knn_exact_grid(d_base_shard, id_base, d_group_queries, K, ids, d, comm, rank, world, base_shards, metric=metric, stream=stream)
    return ids, d


def knn_sharded_torch(d_base_shard: torch.Tensor, id_base: int, d_queries: torch.Tensor, K: int, metric: int = 1, group=None,
                      gather: bool = True, stream: Optional[int] = None):
    """The same decomposition with torch.distributed's all-to-all as the exchange (round 1's path; kept as the yardstick
    tools/bench_knn_sharded.py compares the C-ABI path against)."""
    from . import capi

    if not (d_base_shard.is_cuda and d_queries.is_cuda):
        raise capi.RoarGraphError(capi.RG_ERR_NO_DEVICE, "knn_sharded needs CUDA tensors (there is no CPU fallback)")

    def local_knn(q):
        ids = torch.empty((q.shape[0], K), dtype=torch.int32, device=q.device)
        d = torch.empty((q.shape[0], K), dtype=torch.float32, device=q.device)
        capi.knn_exact_device(d_base_shard, q, K, ids, d, metric=metric, id_base=id_base, stream=stream)
        return ids, d

    def merge(pid, pd):
        G, m, _ = pid.shape
        ids = torch.empty((m, K), dtype=torch.int32, device=pid.device)
        d = torch.empty((m, K), dtype=torch.float32, device=pid.device)
        if m:
            capi.knn_merge_device(pid.contiguous(), pd.contiguous(), ids, d, metric=metric, stream=stream)
        return ids, d

    return knn_sharded_with(local_knn, merge, d_queries, K, group, gather)
