"""Multi-GPU sharding of pose / edge / query batches: one process per GPU, torch.distributed for the plumbing.

The obstacle BVH, robot mesh and node set are replicated on every GPU (a few MB); a batch is split into
``world_size`` contiguous, equally padded index ranges; each rank computes its slice with its own engine handle and
the slices are exchanged with ONE all-gather (NCCL over NVLink on GPUs; gloo in the CPU unit tests).  There is no
other data-path collective: every pose verdict and every query row is independent.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (begin, end, padded_len): contiguous equal slices, the last ones possibly short or empty."""
    per = (n + world - 1) // world if n > 0 else 0
    b = min(rank * per, n)
    e = min(b + per, n)
    return b, e, per


def sharded_rows(n: int, row_shape: Tuple[int, ...], dtype: torch.dtype, device, compute: Callable[[int, int, torch.Tensor], None],
                 group=None, pad_value=0) -> torch.Tensor:
    """Generic "split, compute local slice, all-gather" driver.

    ``compute(begin, end, out)`` fills ``out[: end - begin]`` (a view of this rank's padded slice).  Returns the full
    ``[n, *row_shape]`` tensor on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e, per = shard_bounds(n, rank, world)
    local = torch.full((per, *row_shape), pad_value, dtype=dtype, device=device)
    if e > b:
        compute(b, e, local)
    if world == 1:
        return local[:n]
    full = torch.empty((world * per, *row_shape), dtype=dtype, device=device)
    dist.all_gather_into_tensor(full, local, group=group)
    return full[:n]


def sharded_collide(env, poses: torch.Tensor, group=None) -> torch.Tensor:
    """poses: CUDA tensor [n][6] replicated (or at least valid for this rank's slice) -> uint8 [n] on every rank."""
    n = poses.shape[0]

    def compute(b, e, out):
        env.collide_device(poses[b:e].contiguous(), out=out[: e - b])

    return sharded_rows(n, (), torch.uint8, poses.device, compute, group)


def sharded_edges(env, starts: torch.Tensor, ends: torch.Tensor, sample_dist: float, rot_mode: int, group=None) -> torch.Tensor:
    m = starts.shape[0]

    def compute(b, e, out):
        env.edges_device(starts[b:e].contiguous(), ends[b:e].contiguous(), sample_dist, rot_mode, free_out=out[: e - b])

    return sharded_rows(m, (), torch.uint8, starts.device, compute, group)


def sharded_knn(index, queries: torch.Tensor, k: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Node set replicated, query rows split (the pattern of FLANN's MPI index, lib/flann/src/cpp/flann/mpi/index.h:196-226,
    without its merge step: every rank owns whole rows).  The search kernel writes this rank's (ids, d2) rows straight into
    its slice of the gathered layout and the two all-gathers run in place on those buffers -- no staging copies."""
    nq = queries.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e, per = shard_bounds(nq, rank, world)
    dev = queries.device
    ids = torch.empty((world * per, k), dtype=torch.int32, device=dev)
    d2 = torch.empty((world * per, k), dtype=torch.float32, device=dev)
    lo = rank * per
    if e > b:
        index.knn_device(queries[b:e], k, ids[lo:lo + (e - b)], d2[lo:lo + (e - b)])
    if per > e - b:          # ragged tail: rows that exist only as padding of the equal-sized slices
        ids[lo + (e - b):lo + per].fill_(-1)
        d2[lo + (e - b):lo + per].fill_(float("inf"))
    if world > 1:
        dist.all_gather_into_tensor(ids, ids[lo:lo + per], group=group)
        dist.all_gather_into_tensor(d2, d2[lo:lo + per], group=group)
    return ids[:nq], d2[:nq]


def sharded_radius(index, queries, radius_sq: float, device="cpu", group=None):
    """Radius search with the query rows split over the ranks (node set replicated).  Rows have variable length, so the
    exchange is the counts first, then the packed rows padded to the largest per-rank total (all-gather of the
    counts -> exclusive scan -> padded all-gather of the rows).  ``queries``: numpy / CPU tensor [nq][dim], identical on
    every rank; ``device``: where the collective runs ("cuda" for NCCL, "cpu" for gloo).
    -> (counts int32 [nq], offsets int64 [nq + 1], ids int32 [total], d2 float32 [total]) on every rank, rows in query order,
    each row sorted by (d2, id) -- exactly what ``Index.radiusSearch`` returns for the whole batch."""
    import numpy as np
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
    nq = q.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if nq == 0:   # (the same on every rank: nothing to exchange)
        return (torch.zeros(0, dtype=torch.int32, device=device), torch.zeros(1, dtype=torch.int64, device=device),
                torch.zeros(0, dtype=torch.int32, device=device), torch.zeros(0, dtype=torch.float32, device=device))
    b, e, per = shard_bounds(nq, rank, world)
    if e > b:
        c_loc, _, i_loc, d_loc = index.radiusSearch(q[b:e], radius_sq)
    else:
        c_loc, i_loc, d_loc = np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32)
    if world == 1:
        off = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(c_loc, out=off[1:])
        return (torch.from_numpy(np.ascontiguousarray(c_loc)), torch.from_numpy(off), torch.from_numpy(np.ascontiguousarray(i_loc)),
                torch.from_numpy(np.ascontiguousarray(d_loc)))
    counts_pad = torch.zeros(per, dtype=torch.int32, device=device)
    counts_pad[: e - b] = torch.from_numpy(np.ascontiguousarray(c_loc)).to(device)
    counts_all = torch.empty(world * per, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(counts_all, counts_pad, group=group)
    totals = counts_all.view(world, per).sum(dim=1, dtype=torch.int64)          # rows per rank
    width = int(totals.max().item())
    rows = torch.zeros((2, max(width, 1)), dtype=torch.int32, device=device)      # ids and the bit patterns of d2, one exchange
    if len(i_loc):
        rows[0, : len(i_loc)] = torch.from_numpy(np.ascontiguousarray(i_loc)).to(device)
        rows[1, : len(d_loc)] = torch.from_numpy(np.ascontiguousarray(d_loc).view(np.int32)).to(device)
    rows_all = torch.empty((world * 2, max(width, 1)), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(rows_all, rows, group=group)
    rows_all = rows_all.view(world, 2, max(width, 1))
    tot = [int(t) for t in totals.tolist()]
    ids = torch.cat([rows_all[r, 0, : tot[r]] for r in range(world)])
    d2 = torch.cat([rows_all[r, 1, : tot[r]] for r in range(world)]).view(torch.float32)
    counts = counts_all[:nq]
    offsets = torch.zeros(nq + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts.to(torch.int64), 0, out=offsets[1:])
    return counts, offsets, ids, d2


class PeerGather:
    """Verdict all-gather fused into the collision kernel (include/sffg.h "multi-GPU").

    Every rank owns ``NBUF`` gathered buffers of ``world * per_rank`` verdict bytes plus a few flag words, allocated by the
    engine (plain cudaMalloc, so that they can be exported through CUDA IPC), and maps those of all peers.
    ``collide(env, poses)`` launches the pose kernel with ``world`` destinations -- slice ``rank`` of the current buffer of
    every rank -- so the exchange rides on the kernel's own 32-byte stores over NVLink while it computes; the last CTA to
    finish publishes the call's epoch to every rank, and before touching its destinations a kernel waits until every rank
    has published epoch ``j - NBUF + 2`` (the buffer it is about to overwrite has been consumed everywhere).  With
    NBUF = 4 a rank may run a full kernel ahead of the slowest one: there is no barrier between consecutive calls.
    ``wait(env, step)`` enqueues the wait a consumer of that call's results needs; the contract is that the results of call
    ``s`` are read (enqueued on the same stream) before call ``s + 2`` is enqueued.
    torch.distributed is only the plumbing that ships the 64-byte IPC handles.
    """

    NBUF = 4

    def __init__(self, per_rank: int, group=None):
        import ctypes as C

        from . import _lib
        self._C, self._L, self._check = C, _lib.load(), _lib.check
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerGather supports up to 8 ranks (one NVSwitch domain)")
        self.per = (per_rank + 31) // 32 * 32          # slices start 32-byte aligned
        self.bytes = self.world * self.per
        self.own, handles = [], []
        for b in range(self.NBUF + 1):                  # gathered buffers + one block of flag words
            p = C.c_void_p()
            h = (C.c_uint8 * 64)()
            self._check(self._L.sffg_peer_buffer_create(self.bytes if b < self.NBUF else 256, C.byref(p), h))
            self.own.append(p.value)
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        self.ptrs = []                                  # ptrs[b][r] = buffer b of rank r, valid in this process
        self._opened = []
        for b in range(self.NBUF + 1):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self.own[b])
                else:
                    p = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(everyone[r][b])
                    self._check(self._L.sffg_peer_buffer_open(hb, C.byref(p)))
                    self._opened.append(p.value)
                    row.append(p.value)
            self.ptrs.append(row)
        self._flags = (C.c_void_p * self.world)(*self.ptrs[self.NBUF])      # flag words: [0..7] epochs, [16] CTA counter
        self._done_counter = self.own[self.NBUF] + 64
        self.step = 0
        dist.barrier(group=group)

    def collide(self, env, poses: torch.Tensor) -> int:
        """enqueue the pose kernel on the current stream; verdict i of this rank lands at offset rank*per + i of the current
        gathered buffer of EVERY rank.  Returns the step number to pass to :meth:`wait` / :meth:`view`."""
        n = poses.shape[0]
        assert n <= self.per and poses.is_cuda and poses.is_contiguous()
        j = self.step
        self.step += 1
        b = j % self.NBUF
        dests = (self._C.c_void_p * self.world)(*[self.ptrs[b][r] + self.rank * self.per for r in range(self.world)])
        st = torch.cuda.current_stream(poses.device).cuda_stream
        epoch = j + 1
        wait = max(0, epoch - self.NBUF + 2)
        self._check(self._L.sffg_collide_poses_gather_sync_device(env._h, poses.data_ptr(), int(poses.dtype == torch.float64), n, dests,
                                                                  self._flags, self.world, self.rank, epoch, wait, self._done_counter, st))
        return j

    def wait(self, env, step: int, device=None) -> None:
        """enqueue, on the current stream, the wait for every rank's results of call ``step``"""
        st = torch.cuda.current_stream(device).cuda_stream
        self._check(self._L.sffg_peer_wait_device(env._h, self._flags, self.world, self.rank, step + 1, st))

    def view(self, step: int, device) -> torch.Tensor:
        """the gathered buffer of call ``step`` on this rank as a torch uint8 tensor [world, per] (no copy)"""
        ptr = self.own[step % self.NBUF]
        iface = {"shape": (self.world, self.per), "typestr": "|u1", "data": (ptr, False), "version": 2}
        holder = type("_Dev", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=device)

    def close(self) -> None:
        for p in self._opened:
            self._L.sffg_peer_buffer_close(p)
        self._opened = []
        for p in self.own:
            self._L.sffg_peer_buffer_destroy(p)
        self.own = []


class PeerRows:
    """k-NN row all-gather fused into the search kernels (include/sffg.h, sffg_knn_gather_device).

    Every rank owns two sets of gathered buffers -- ids int32 [world * per][k] and d2 float32 [world * per][k] -- allocated
    by the engine and exported through CUDA IPC.  ``knn(index, queries_local, k)`` launches the search with ``world``
    destinations (this rank's row range in every rank's current buffers), then a signal + wait barrier over the flag words:
    when it has run, every rank holds every rank's rows.  The two buffer sets alternate, so a rank may already write call
    j + 1 while a slower rank still reads call j; the contract is that the rows of call j are read (on the same stream)
    before call j + 2 is enqueued.  torch.distributed only ships the IPC handles.
    """

    NBUF = 2

    def __init__(self, per_rank_rows: int, k: int, group=None):
        import ctypes as C

        from . import _lib
        self._C, self._L, self._check = C, _lib.load(), _lib.check
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerRows supports up to 8 ranks (one NVSwitch domain)")
        self.per, self.k = per_rank_rows, k
        self.bytes = self.world * self.per * k * 4
        self.own, handles = [], []
        for b in range(2 * self.NBUF + 1):              # (ids, d2) x NBUF + one block of flag words
            p = C.c_void_p()
            h = (C.c_uint8 * 64)()
            self._check(self._L.sffg_peer_buffer_create(self.bytes if b < 2 * self.NBUF else 256, C.byref(p), h))
            self.own.append(p.value)
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        self.ptrs, self._opened = [], []
        for b in range(2 * self.NBUF + 1):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self.own[b])
                else:
                    p = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(everyone[r][b])
                    self._check(self._L.sffg_peer_buffer_open(hb, C.byref(p)))
                    self._opened.append(p.value)
                    row.append(p.value)
            self.ptrs.append(row)
        self._flags = (C.c_void_p * self.world)(*self.ptrs[2 * self.NBUF])
        self.step = 0
        dist.barrier(group=group)

    def knn(self, index, queries_local: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (ids [world * per, k], d2 [world * per, k]) views of this rank's gathered buffers of this call; rank r's rows
        start at r * per.  Enqueued on the current stream (search, then the barrier)."""
        n = queries_local.shape[0]
        assert k == self.k and n <= self.per and queries_local.is_cuda and queries_local.is_contiguous()
        j = self.step
        self.step += 1
        b = j % self.NBUF
        off = self.rank * self.per * k * 4
        ids_d = (self._C.c_void_p * self.world)(*[self.ptrs[2 * b][r] + off for r in range(self.world)])
        d2_d = (self._C.c_void_p * self.world)(*[self.ptrs[2 * b + 1][r] + off for r in range(self.world)])
        st = torch.cuda.current_stream(queries_local.device).cuda_stream
        self._check(self._L.sffg_knn_gather_device(index._h, queries_local.data_ptr(), n, k, ids_d, d2_d, self.world, st))
        self._check(self._L.sffg_peer_barrier_device(None, self._flags, self.world, self.rank, j + 1, st))
        return self._view(self.own[2 * b], torch.int32, queries_local.device), self._view(self.own[2 * b + 1], torch.float32, queries_local.device)

    def _view(self, ptr, dtype, device):
        typestr = "<i4" if dtype == torch.int32 else "<f4"
        iface = {"shape": (self.world * self.per, self.k), "typestr": typestr, "data": (ptr, False), "version": 2}
        holder = type("_Dev", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=device)

    def close(self) -> None:
        for p in self._opened:
            self._L.sffg_peer_buffer_close(p)
        self._opened = []
        for p in self.own:
            self._L.sffg_peer_buffer_destroy(p)
        self.own = []


def gathered_collide(env, pg: "PeerGather", poses_local: torch.Tensor) -> torch.Tensor:
    """one fused step: local slice -> every rank's gathered buffer, then wait for everyone's slice; returns [world, per]"""
    j = pg.collide(env, poses_local)
    pg.wait(env, j, poses_local.device)
    return pg.view(j, poses_local.device)
