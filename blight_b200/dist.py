"""Multi-GPU layer of the query path: one process per GPU, torch.distributed for the plumbing.

ReplicaSet       small / medium indices: the whole index on every GPU, reads split in contiguous chunks, no collective
                 on the data path (the three counters are summed once at the end).
PartitionedSet   large indices: the 2^n MPHF groups are cut into `world` contiguous ranges balanced by k-mer count; a
                 rank holds only its slice (blight_flat_slice). Per batch: front end on the rank that holds the reads
                 -> bin (canon, minimizer) by owner -> ONE all-to-all out -> lookup at the owner -> ONE all-to-all back
                 -> scatter into read order. Ids are global in both modes.

PartitionedSet has two data paths. `query_reads_fused` is the product on NVLink boxes: the exchange is fused into the
two kernels either side of it (csrc/part_kernels.cu) — the source stores one 32-byte record per super-k-mer straight into
the owner's inbox (peer memory), the owner stores its ids straight into the source's memory (contiguous 32-bit streams the
source then widens into read order, or — return_path "direct" — int64 ids in their final place); the ordering between GPUs
is device-side flags in peer memory (csrc/part_session.cu), so the only collectives of a batch are the agreement on the
number of sub-batches and one all-reduce of the counters. `enable_fused(mode="legacy")` keeps round 1's Python pipeline
(a counter all-to-all per sub-batch) for comparison. `query_reads` is the plain
formulation (NCCL all-to-all of (canon, minimizer) out and ids back), kept as the fallback when an inbox overflows and as
the path the CPU `gloo` tests drive.

The routing (`PartitionPlan`, `exchange_lookup`) is device-agnostic torch code so that world_size-2 `gloo` tests on CPU
exercise exactly the logic the NCCL path runs; on CUDA the binning and the final scatter are the library's own kernels
(blight_owner_count / blight_owner_scatter / blight_scatter_ids), never a torch sort.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import api


# base positions per sub-batch of the fused path when the caller does not say: few, large sub-batches (fewer kernel ends and
# flag rounds). On 8 B200s, 480 M k-mers per GPU: counting 15.2 / 14.6 / 14.3 / 14.2 ms at 32 / 64 / 128 / 256 M; ids on the
# default direct return path 19.5 / 18.4 / 18.0 ms at 32 / 64 / 128 M (profiles/r02_v18_bench_n8.json)
DEFAULT_SUB_COUNTING = 256 << 20
DEFAULT_SUB_IDS = 128 << 20
DEFAULT_ORDER = "serial"  # kernel order of the fused partitioned path when neither the caller nor BLIGHT_PART_ORDER says (part_session.cu)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class PartitionPlan:
    """Contiguous ranges of MPHF groups per rank. cuts has world+1 ascending entries, cuts[0]=0, cuts[-1]=n_mphf."""
    cuts: List[int]
    lb: int  # log2(buckets per MPHF group): group = minimizer >> lb (blight.cpp:722)

    @property
    def world(self) -> int:
        return len(self.cuts) - 1

    @staticmethod
    def balanced(group_sizes: np.ndarray, world: int, lb: int) -> "PartitionPlan":
        """Cuts after the group at which the running k-mer count first reaches r/world of the total (every rank gets
        at least one group when there are enough groups)."""
        g = np.asarray(group_sizes, dtype=np.float64)
        n = len(g)
        if world > n:
            raise ValueError(f"cannot partition {n} MPHF groups over {world} ranks (use n >= log2(world))")
        csum = np.cumsum(g)
        total = csum[-1] if n else 0.0
        cuts = [0]
        for r in range(1, world):
            c = int(np.searchsorted(csum, total * r / world, side="left")) + 1
            c = max(c, cuts[-1] + 1)          # at least one group per rank
            c = min(c, n - (world - r))       # leave one group for each remaining rank
            cuts.append(c)
        cuts.append(n)
        return PartitionPlan(cuts, lb)

    def owner_of(self, mini: torch.Tensor) -> torch.Tensor:
        """Owner rank of each minimizer (any device)."""
        g = (mini.to(torch.int64) & 0xFFFFFFFF) >> self.lb
        inner = torch.tensor(self.cuts[1:-1], dtype=torch.int64, device=mini.device)
        return torch.searchsorted(inner, g, right=True)

    def group_range(self, rank: int) -> Tuple[int, int]:
        return self.cuts[rank], self.cuts[rank + 1]


def _torch_bin(canon: torch.Tensor, mini: torch.Tensor, plan: PartitionPlan):
    """Device-agnostic binning (used on CPU by the gloo tests): stable sort by owner."""
    owner = plan.owner_of(mini)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=plan.world).to(torch.int64)
    return canon[order], mini[order], order, counts


def _cuda_bin(canon: torch.Tensor, mini: torch.Tensor, plan: PartitionPlan):
    """Binning with the library's kernels: count, prefix on the device, warp-aggregated scatter."""
    L = api.lib()
    n = canon.numel()
    dev = canon.device
    st = torch.cuda.current_stream().cuda_stream
    cuts = torch.tensor(plan.cuts, dtype=torch.int32, device=dev)
    counts = torch.zeros(plan.world, dtype=torch.int64, device=dev)
    api._check(L.blight_owner_count(mini.data_ptr(), n, cuts.data_ptr(), plan.world, plan.lb, counts.data_ptr(), st))
    cursors = torch.cumsum(counts, 0) - counts
    send_canon = torch.empty_like(canon)
    send_mini = torch.empty_like(mini)
    send_src = torch.empty(n, dtype=torch.int64, device=dev)
    api._check(L.blight_owner_scatter(canon.data_ptr(), mini.data_ptr(), n, cuts.data_ptr(), plan.world, plan.lb,
                                      cursors.data_ptr(), send_canon.data_ptr(), send_mini.data_ptr(), send_src.data_ptr(), st))
    return send_canon, send_mini, send_src, counts


def _cuda_scatter(ids_back: torch.Tensor, src: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.int64, device=ids_back.device)
    api._check(api.lib().blight_scatter_ids(ids_back.data_ptr(), src.data_ptr(), n, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    return out


def exchange_lookup(canon: torch.Tensor, mini: torch.Tensor, plan: PartitionPlan,
                    lookup_local: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], group=None) -> torch.Tensor:
    """ids for (canon, mini) pairs when every rank holds only its group range.
    lookup_local(canon, mini) -> int64 ids answers pairs whose groups this rank owns."""
    world = dist.get_world_size(group)
    assert world == plan.world
    n = canon.numel()
    on_cuda = canon.is_cuda
    send_canon, send_mini, src, counts = (_cuda_bin if on_cuda else _torch_bin)(canon.contiguous(), mini.contiguous(), plan)
    # split sizes: one small all-to-all of the counts, then the payloads
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_splits = counts.tolist()
    out_splits = recv_counts.tolist()
    m = int(sum(out_splits))
    recv_canon = torch.empty(m, dtype=canon.dtype, device=canon.device)
    recv_mini = torch.empty(m, dtype=mini.dtype, device=mini.device)
    dist.all_to_all_single(recv_canon, send_canon, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    dist.all_to_all_single(recv_mini, send_mini, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    ids_remote = lookup_local(recv_canon, recv_mini)
    ids_back = torch.empty(n, dtype=torch.int64, device=canon.device)
    dist.all_to_all_single(ids_back, ids_remote.contiguous(), output_split_sizes=in_splits, input_split_sizes=out_splits, group=group)
    if on_cuda:
        return _cuda_scatter(ids_back, src, n)
    out = torch.empty(n, dtype=torch.int64, device=canon.device)
    out[src] = ids_back
    return out


def all_reduce_counters(ctr: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of [found, not_found, queries, invalid] over ranks (the only collective of replica mode)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ctr, op=dist.ReduceOp.SUM, group=group)
    return ctr


class ReplicaSet:
    """Whole index on every GPU; each rank queries its contiguous chunk of the reads."""

    def __init__(self, flat: api.FlatIndex, device: int, group=None):
        self.group = group
        self.index = flat.upload(device)
        self.k = self.index.k

    def query_reads_sharded(self, bases: torch.Tensor, read_off: torch.Tensor, want_ids: bool = True):
        """bases / read_off describe the WHOLE batch (same on every rank, on this rank's device); returns this rank's
        (ids of its chunk or None, global counters)."""
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        n_reads = read_off.numel() - 1
        lo, hi = shard_range(n_reads, rank, world)
        off = read_off[lo:hi + 1].contiguous()
        b0, b1 = int(off[0]), int(off[-1])
        sub = bases[b0:b1]
        if (sub.data_ptr() & 15) != 0:
            sub = sub.clone()  # keep the 16-byte aligned fast path of the front end
        off = off - b0
        lens = off[1:] - off[:-1]
        nk = torch.clamp(lens - (self.k - 1), min=0)
        koff = torch.zeros(hi - lo + 1, dtype=torch.int64, device=bases.device)
        torch.cumsum(nk, 0, out=koff[1:])
        total = int(koff[-1])
        ids, ctr = self.index.query_reads(sub, off, koff, total, want_ids=want_ids)
        return (ids[:total] if want_ids else None), all_reduce_counters(ctr, self.group)


class PartitionedSet:
    """Minimizer-bucket partition: this rank holds MPHF groups [cuts[rank], cuts[rank+1])."""

    def __init__(self, plan: PartitionPlan, local_flat: api.FlatIndex, device: int, k: int, m: int, group=None):
        self.plan, self.group, self.k, self.m = plan, group, k, m
        self.index = local_flat.upload(device)

    @classmethod
    def from_full(cls, flat: Optional[api.FlatIndex], device: int, workdir: str, group=None) -> "PartitionedSet":
        """Rank 0 holds the full flat index: it plans, slices and saves one blob per rank; every rank loads its own."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        meta = [None]
        if rank == 0:
            info = flat.info()
            lb = 2 * info["m"] - 1 - info["n_log2"]
            plan = PartitionPlan.balanced(flat.group_sizes(), world, lb)
            for r in range(world):
                a, b = plan.group_range(r)
                flat.slice(a, b).save(os.path.join(workdir, f"part{r}.blflat"))
            meta = [(plan.cuts, lb, info["k"], info["m"])]
        dist.broadcast_object_list(meta, src=0, group=group)
        cuts, lb, k, m = meta[0]
        local = api.FlatIndex.load(os.path.join(workdir, f"part{rank}.blflat"))
        return cls(PartitionPlan(list(cuts), lb), local, device, k, m, group)

    # ---- fused path: peer-memory stores inside the kernels -------------------------------------------------------
    def disable_fused(self):
        """Unmaps the peers' buffers, then (after a barrier) frees this rank's."""
        if getattr(self, "_session", None) is not None:
            torch.cuda.synchronize()
            if self._world > 1:
                dist.barrier(group=self.group)  # nobody may still be storing into this rank's buffers
            self._session.close()
            self._session = None
            self._ids_view = None
            return
        if not hasattr(self, "_inbox"):
            return
        torch.cuda.synchronize()
        for pb in self._peers:
            pb.close()
        self._peers = []
        if self._world > 1:
            dist.barrier(group=self.group)
        self._inbox.close()
        if self._ret:
            self._ret.close()
        del self._inbox, self._ret, self._side

    def enable_fused(self, want_ids: bool = True, sub_positions: Optional[int] = None, records_per_position: Optional[float] = None,
                     ids_capacity: int = 0, mode: Optional[str] = None, order: Optional[str] = None, return_path: Optional[str] = None):
        """Allocates and exchanges the peer buffers. mode "session" (default): csrc/part_session.cu — per rank a double
        buffered inbox of world regions of `cap` records (written by the sources), a mailbox of flags, and for the id mode
        an id array of `ids_capacity` int64 (grown on demand by query_reads_fused) plus, on the default "stream" return_path,
        return regions of 32-bit ids and a side table ("direct": the owners store int64 ids straight into the id array).
        mode "legacy": round 1's Python pipeline (one NCCL all-to-all of the record counts per sub-batch), kept for comparison.
        sub_positions = base positions per sub-batch (one dispatch and one lookup kernel each)."""
        self.disable_fused()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        dev = torch.device("cuda", self.index.device)
        if sub_positions is None:
            sub_positions = DEFAULT_SUB_IDS if want_ids else DEFAULT_SUB_COUNTING
        mode = mode or os.environ.get("BLIGHT_PART_PIPELINE", "session")
        if mode != "legacy":
            if records_per_position is None:
                records_per_position = min(0.25, max(0.03, 0.5 / world))
            max_cap = (1 << 24) - 1
            sub_positions = min(int(sub_positions), int(max_cap / records_per_position))
            self._sub = max(256, (sub_positions // 256) * 256)
            self._cap = min(max(1024, int(self._sub * records_per_position)), max_cap)
            self._world, self._rank, self._fused_ids = world, rank, want_ids
            self._session_args = (world, rank)
            self._order = order or os.environ.get("BLIGHT_PART_ORDER") or DEFAULT_ORDER  # "serial" | "ahead" | "overlap"
            self._return_path = return_path  # None: BLIGHT_PART_RETURN or "stream"
            self._make_session(int(ids_capacity) if want_ids else 0)
            return
        if records_per_position is None:
            # a read batch has ~0.07 super-k-mers per base, spread over `world` owners; an overflow is detected and the
            # batch answered through the plain path, so this only has to be generous, not safe
            records_per_position = min(0.25, max(0.03, 0.5 / world))
        max_cap = (1 << 24) - 1  # the packed (slots, k-mers) counter keeps 24 bits of slots
        sub_positions = min(int(sub_positions), int(max_cap / records_per_position))
        self._sub = max(256, (sub_positions // 256) * 256)
        self._cap = min(max(1024, int(self._sub * records_per_position)), max_cap)
        self._kcap = self._sub
        self._world, self._rank, self._fused_ids = world, rank, want_ids
        region_bytes = self._cap * api.RUN_RECORD_BYTES
        ret_bytes = self._kcap * 4
        self._inbox = api.PeerBuffer.alloc(2 * world * region_bytes)
        self._ret = api.PeerBuffer.alloc(2 * world * ret_bytes) if want_ids else None
        self._side = torch.empty(2 * world * self._cap * 16, dtype=torch.uint8, device=dev) if want_ids else None
        mine = (self._inbox.handle, self._ret.handle if self._ret else b"")
        everyone = [None] * world
        if world > 1:
            dist.all_gather_object(everyone, mine, group=self.group)
        else:
            everyone = [mine]
        self._peers = []  # keep the mappings alive
        inbox_ptr, ret_ptr = [0] * world, [0] * world
        for r, (hi, hr) in enumerate(everyone):
            if r == rank:
                inbox_ptr[r], ret_ptr[r] = self._inbox.ptr, (self._ret.ptr if self._ret else 0)
                continue
            pb = api.PeerBuffer.open(hi, 2 * world * region_bytes)
            self._peers.append(pb)
            inbox_ptr[r] = pb.ptr
            if hr:
                pr = api.PeerBuffer.open(hr, 2 * world * ret_bytes)
                self._peers.append(pr)
                ret_ptr[r] = pr.ptr
        self._routes, self._regions, self._ret_at, self._ret_mine, self._side_at = [], [], [], [], []
        for b in range(2):
            rt = api.PartRoute()
            rt.world, rt.rank, rt.lb, rt.cap, rt.kcap = world, rank, self.plan.lb, self._cap, self._kcap
            for i, c in enumerate(self.plan.cuts):
                rt.cuts[i] = c
            for d in range(world):  # my region in owner d's inbox: [buffer b][source = me]
                rt.inbox[d] = inbox_ptr[d] + (b * world + rank) * region_bytes
            side_b = (self._side.data_ptr() + b * world * self._cap * 16) if want_ids else 0
            rt.side = side_b
            self._routes.append(rt)
            self._side_at.append(side_b)
            self._regions.append([self._inbox.ptr + (b * world + s_) * region_bytes for s_ in range(world)])
            # where I, as an owner, return ids to source s: [buffer b][owner = me] of s's return area; and my own area
            self._ret_at.append([ret_ptr[s_] + (b * world + rank) * ret_bytes for s_ in range(world)] if want_ids else None)
            self._ret_mine.append((self._ret.ptr + b * world * ret_bytes) if want_ids else 0)
        self._counts = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(2)]
        self._recv_counts = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(2)]
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        # owners return slice-local 32-bit ids; the scatter adds the owner's first identifier back
        bases = [None] * world
        if world > 1:
            dist.all_gather_object(bases, int(self.index.info["id_base"]), group=self.group)
        else:
            bases = [int(self.index.info["id_base"])]
        self._id_bases = np.asarray(bases, dtype=np.uint64)
        self._side_stream = torch.cuda.Stream(device=dev)
        self._ev_pub = torch.cuda.Event()
        self._ev_scat = [torch.cuda.Event(), torch.cuda.Event()]
        if world > 1:
            dist.barrier(group=self.group)

    def _make_session(self, ids_capacity: int):
        """(Re)creates this rank's session with an id array of ids_capacity entries and connects the peers (collective)."""
        world, rank = self._session_args
        if getattr(self, "_session", None) is not None:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier(group=self.group)
            self._session.close()
        dev = torch.device("cuda", self.index.device)
        # a sub-batch spreads its k-mers over `world` owners: return regions of 2.5 / world of it (an overflow is detected)
        ret_kmers = self._sub if world <= 2 else int(self._sub * 2.5 / world)
        self._session = api.PartSession(self.index, world, rank, self.plan.lb, self.plan.cuts, self._sub, self._cap, ids_capacity,
                                        order=self._order, return_path=self._return_path, ret_kmers=ret_kmers)
        everyone = [None] * world
        mine = (self._session.handles(), ids_capacity, int(self.index.info["id_base"]))
        if world > 1:
            dist.all_gather_object(everyone, mine, group=self.group)
        else:
            everyone = [mine]
        for r, (h, cap_r, base_r) in enumerate(everyone):
            if r != rank:
                self._session.connect_ipc(r, h, cap_r, base_r)
        self._ids_view = self._session.ids_tensor(dev)
        self._err_seen = 0
        if world > 1:
            dist.barrier(group=self.group)

    def overflowed(self) -> bool:
        """True if the last query_reads_fused on this rank dropped records (only meaningful with check_overflow=False)."""
        if getattr(self, "_session", None) is not None:
            self._err_seen |= self._session.status(reset=True)
            return bool(self._err_seen & api.PART_OVERFLOW)
        return bool(int(self._err.item()))

    def query_reads_fused(self, bases: torch.Tensor, read_off: torch.Tensor, kmer_off: Optional[torch.Tensor] = None,
                          total_kmers: int = 0, want_ids: bool = True, ids: Optional[torch.Tensor] = None, check_overflow: bool = True):
        """Reads held by THIS rank -> (ids in read order or None, GLOBAL counters [found, not_found, queries, invalid] summed
        over all ranks). Collective: every rank of the group must call it, with the same want_ids. Per sub-batch i, on the
        current stream: dispatch(i) -> all-to-all of the counters (publishes the records of i and, because every owner
        finished lookup(i-1) before entering it, the ids of i-1) -> scatter(i-1) -> lookup(i)."""
        if getattr(self, "_session", None) is not None:
            return self._query_session(bases, read_off, kmer_off, total_kmers, want_ids, ids, check_overflow)
        if not hasattr(self, "_inbox"):
            raise RuntimeError("call enable_fused() first")
        world = self._world
        dev = bases.device
        if want_ids and not self._fused_ids:
            raise ValueError("enable_fused(want_ids=False) was asked for")
        if want_ids and ids is None:
            ids = torch.empty(max(int(total_kmers), 1), dtype=torch.int64, device=dev)
        if want_ids:
            ids.record_stream(self._side_stream)
        ctr = torch.zeros(api.N_CTR, dtype=torch.int64, device=dev)
        total = bases.numel()
        n_sub = (total + self._sub - 1) // self._sub
        if world > 1:
            t = torch.tensor([n_sub], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            n_sub = int(t.item())
        self._err.zero_()
        koff = kmer_off if want_ids else None
        max_rec = world * self._cap

        main = torch.cuda.current_stream()
        side = self._side_stream

        def scatter(b):
            # on a second stream, next to the lookup of the following sub-batch (it only moves bytes)
            self._ev_pub.record(main)
            side.wait_event(self._ev_pub)
            api.part_scatter(self._side_at[b], self._cap, self._counts[b], self._ret_mine[b], self._kcap, world, max_rec, ids, self._id_bases, stream=side)
            self._ev_scat[b].record(side)

        pending = [False, False]

        def dispatch(i):
            b = i & 1
            if pending[b]:
                main.wait_event(self._ev_scat[b])  # the scatter of sub-batch i-2 still reads this buffer's counters and side table
                pending[b] = False
            self._counts[b].zero_()
            if i * self._sub < total:
                api.part_dispatch(self.k, self.m, bases, read_off, koff, self._routes[b], self._counts[b], ctr, self._err,
                                  i * self._sub, min(total, (i + 1) * self._sub))

        def exchange(i):
            b = i & 1
            if world > 1:
                dist.all_to_all_single(self._recv_counts[b], self._counts[b], group=self.group)
                return self._recv_counts[b]
            return self._counts[b]

        def lookup(i, rcv):
            b = i & 1
            api.part_lookup(self.index, self._regions[b], rcv, self._ret_at[b] if want_ids else None, self._cap, self._kcap, ctr)

        if os.environ.get("BLIGHT_PART_ORDER") == "ahead":
            # experimental order (not measured at N = 8 yet): dispatch(i+1) is issued before lookup(i), so by the time the
            # counter exchange of i+1 is entered every rank's dispatch is long done and the barrier only sees lookup skew
            rcv = None
            if n_sub > 0:
                dispatch(0)
                rcv = exchange(0)
            for i in range(n_sub):
                if want_ids and i > 0:
                    scatter((i - 1) & 1)
                    pending[(i - 1) & 1] = True
                if i + 1 < n_sub:
                    dispatch(i + 1)
                lookup(i, rcv)
                if i + 1 < n_sub:
                    rcv = exchange(i + 1)
        else:
            for i in range(n_sub):
                dispatch(i)
                rcv = exchange(i)
                if want_ids and i > 0:
                    scatter((i - 1) & 1)
                    pending[(i - 1) & 1] = True
                lookup(i, rcv)
        e = self._err.to(torch.int64)
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(ctr, group=self.group)  # also the barrier after which the last ids have landed
        if want_ids and n_sub > 0:
            scatter((n_sub - 1) & 1)
            pending[(n_sub - 1) & 1] = True
        for b in range(2):
            if pending[b]:
                main.wait_event(self._ev_scat[b])
        if check_overflow:
            host = torch.cat([e, ctr]).cpu()
            if int(host[1 + api.CTR_INVALID]):
                raise api.InvalidBase(api.ERR_INVALID_BASE, "Invalid char in DNA")  # std::domain_error, kmer.h:68
            if int(host[0]):
                # a region overflowed (far more super-k-mers per base than a read batch has): the plain paths have no such limit
                if world == 1:
                    return self.index.query_reads(bases, read_off, kmer_off, total_kmers, want_ids=want_ids)
                ids, c = self.query_reads(bases, read_off, kmer_off, total_kmers)
                return (ids if want_ids else None), all_reduce_counters(c, self.group)
        return (ids[:total_kmers] if want_ids else None), ctr

    def _query_session(self, bases, read_off, kmer_off, total_kmers, want_ids, ids, check_overflow):
        """query_reads_fused through csrc/part_session.cu. With want_ids the result is a view of the session's id array
        (valid until the next call) unless the caller passed `ids`, which then receives a copy."""
        world = self._world
        dev = bases.device
        if want_ids and not self._fused_ids:
            raise ValueError("enable_fused(want_ids=False) was asked for")
        total = bases.numel()
        n_sub = self._session.sub_batches(total, want_ids)
        need = int(total_kmers) if want_ids else 0
        if world > 1:
            t = torch.tensor([n_sub, need], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            n_sub, need_all = int(t[0].item()), int(t[1].item())
        else:
            need_all = need
        if want_ids and need_all > self._session.ids_capacity:
            self._make_session(need_all + need_all // 8)  # collective: every rank sees the same maximum
        ctr = torch.zeros(api.N_CTR, dtype=torch.int64, device=dev)
        self._err_seen = 0
        self._session.query(bases, read_off, kmer_off if want_ids else None, n_sub, ctr)
        if world > 1:
            dist.all_reduce(ctr, group=self.group)
        out = None
        if want_ids:
            out = self._ids_view[:total_kmers]
            if ids is not None:
                ids[:total_kmers].copy_(out)
                out = ids[:total_kmers]
        if check_overflow:
            flags = torch.tensor([self._session.status(reset=True)], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
            flags = int(flags.item())
            if int(ctr[api.CTR_INVALID].item()):
                raise api.InvalidBase(api.ERR_INVALID_BASE, "Invalid char in DNA")  # std::domain_error, kmer.h:68
            if flags & api.PART_TIMEOUT:
                raise api.BlightError(api.ERR_CUDA, "partitioned query: a peer's flag never arrived")
            if flags & api.PART_OVERFLOW:
                # a region overflowed (far more super-k-mers per base than a read batch has): the plain paths have no such limit
                if world == 1:
                    return self.index.query_reads(bases, read_off, kmer_off, total_kmers, want_ids=want_ids)
                ids2, c = self.query_reads(bases, read_off, kmer_off, total_kmers)
                return (ids2 if want_ids else None), all_reduce_counters(c, self.group)
        return out, ctr

    def query_kmers(self, canon: torch.Tensor, mini: torch.Tensor) -> torch.Tensor:
        return exchange_lookup(canon, mini, self.plan, lambda c, mn: self.index.query_kmers(c, mini=mn), self.group)

    def query_reads(self, bases: torch.Tensor, read_off: torch.Tensor, kmer_off: torch.Tensor, total_kmers: int):
        """Reads held by THIS rank -> (ids in read order, local counters [found, not_found, queries, invalid])."""
        canon, mini, fctr = api.reads_to_kmers(self.k, self.m, bases, read_off, kmer_off, total_kmers)
        ids = self.query_kmers(canon.contiguous(), mini.contiguous())
        found = (ids >= 0).sum()
        ctr = torch.stack([found, ids.numel() - found, torch.tensor(ids.numel(), device=ids.device), fctr[api.CTR_INVALID]]).to(torch.int64)
        return ids, ctr
