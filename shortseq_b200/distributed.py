"""Multi-GPU dedup counting: local tables merged by a hash-partitioned all-to-all.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Each rank packs and
counts its own shard of the reads with no communication; the only exchange step moves
UNIQUES, not reads:

  1. export the local table grouped by owner rank (owner = top log2(P) bits of the key hash;
     the table is hash-ordered, so the export kernel writes the P send segments directly),
  2. all-to-all of the segment sizes, then all-to-all-v of (words, lens, counts),
  3. every rank adds the tuples it received to its owner table (weighted insert).

The global counter is the disjoint union of the P owner tables.  The reference has no
multi-process mode; this is new.
"""
import os
import time

import torch
import torch.distributed as dist

from ._lib import CLASS_64

_TIMING = os.environ.get("SSQ_MERGE_TIMING") == "1"      # development: print the wall time of each merge phase (adds syncs)


def _tick(label, t0, ctx):
    if not _TIMING:
        return t0
    torch.cuda.synchronize(ctx.device)
    t1 = time.perf_counter()
    if dist.get_rank() == 0:
        print(f"  [merge] {label}: {(t1 - t0) * 1e3:.2f} ms", flush=True)
    return t1


def exchange_counts(send_counts, group=None):
    """send_counts[p] = tuples this rank sends to rank p  ->  recv_counts[p] = tuples rank p sends here."""
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def exchange_tuples(words, lens, counts, send_counts, recv_counts, group=None):
    """all-to-all-v of the exported tuples.  words [n] or [n, 3] int64, lens uint8, counts int64, all laid
    out as P consecutive segments of send_counts[p] tuples.  Works on any backend (NCCL on GPUs, gloo on CPU)."""
    send = [int(x) for x in send_counts.tolist()]
    recv = [int(x) for x in recv_counts.tolist()]
    total = sum(recv)
    out_words = words.new_empty((total,) + tuple(words.shape[1:]))
    out_lens = lens.new_empty((total,))
    out_counts = counts.new_empty((total,))
    dist.all_to_all_single(out_words, words.contiguous(), recv, send, group=group)
    dist.all_to_all_single(out_lens, lens.contiguous(), recv, send, group=group)
    dist.all_to_all_single(out_counts, counts.contiguous(), recv, send, group=group)
    return out_words, out_lens, out_counts


def merge_alltoall(local, group=None, owner=None):
    """Merge every rank's DeviceCounter `local` into per-rank owner tables.

    -> the DeviceCounter holding the keys this rank owns (global counts).  `owner` may be a
    pre-sized DeviceCounter created with hash_rot = log2(world) to reuse across calls.
    """
    from .counter import DeviceCounter
    world = dist.get_world_size(group)
    if world & (world - 1):
        raise ValueError("merge_alltoall needs a power-of-two world size")
    rot = world.bit_length() - 1
    t0 = _tick("start", time.perf_counter(), local.ctx) if _TIMING else 0.0
    keys, counts, _, parts = local.export(world)
    t0 = _tick("export", t0, local.ctx)
    if world == 1:
        recv_w, recv_l, recv_c = keys.words, keys.lens, counts
    else:
        recv_counts = exchange_counts(parts, group)
        t0 = _tick("exchange sizes", t0, local.ctx)
        recv_w, recv_l, recv_c = exchange_tuples(keys.words, keys.lens, counts, parts, recv_counts, group)
        t0 = _tick("exchange tuples", t0, local.ctx)
    if owner is None:
        owner = DeviceCounter(local.klass, expected_unique=int(recv_l.numel()), hash_rot=rot, device=local.ctx.device)
    owner.merge(recv_w, recv_l, recv_c)
    _tick("merge into owner table", t0, local.ctx)
    return owner


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank cannot share or map the peer buffers (CUDA IPC); callers fall back to
    merge_alltoall (NCCL)."""


class PeerExchange:
    """Receive buffers of this rank for the multi-GPU merge, mapped into every other rank of the box through CUDA
    IPC, so that ssq_counter_export_to on the sending GPU stores the tuples an owner is due straight into the
    owner's memory over NVLink -- the export kernel is the exchange; NCCL only carries the small size matrix.
    One object per process, reused across merges (buffers grow collectively when a merge needs more)."""

    def __init__(self, ctx, group=None):
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap = 0                                # tuples the receive buffers hold
        self.rb_cap = 0                             # sender regions per owner the region-base buffer holds
        self.mine = [0, 0, 0, 0]                    # words / lens / counts / region-base receive buffers (device pointers)
        self.peer = [[0] * self.world for _ in range(4)]
        self._opened = []

    def _release(self):
        """Collective.  CUDA requires every importer to close its mapping (cudaIpcCloseMemHandle) BEFORE the exporter
        frees the memory, so: (1) unmap the peers' buffers, (2) barrier, (3) free this rank's own."""
        from . import _lib
        lib, h = _lib.lib(), self.ctx.bind()
        for p in self._opened:
            lib.ssq_ipc_close(h, p)
        self._opened = []
        self.peer = [[0] * self.world for _ in range(4)]
        dist.barrier(self.group)                   # nobody still maps what is freed next
        for p in self.mine:
            if p:
                lib.ssq_free(h, p)
        self.mine = [0, 0, 0, 0]
        self.cap = 0
        self.rb_cap = 0

    def close(self):
        """Collective: unmap the peers' buffers, barrier, then free this rank's."""
        torch.cuda.synchronize(self.ctx.device)
        dist.barrier(self.group)                   # nobody is still writing into, or reading from, the buffers
        self._release()

    def ensure(self, need, need_regions=0):
        """Collective (every rank passes the same arguments): receive buffers for at least `need` tuples and, per
        sending rank, the offsets of `need_regions` sender regions."""
        if need <= self.cap and need_regions <= self.rb_cap:
            return
        import ctypes as C
        from . import _lib
        lib, h = _lib.lib(), self.ctx.bind()
        torch.cuda.synchronize(self.ctx.device)
        dist.barrier(self.group)                   # nobody is still writing into, or reading from, the old buffers
        self._release()
        cap = int(need * 1.25) + 1024
        rb_cap = max(int(need_regions), 1)
        handles = torch.zeros(4 * 64, dtype=torch.uint8)
        failure = None
        try:
            for k, nbytes in enumerate((8 * cap, cap, 8 * cap, 8 * self.world * (rb_cap + 1))):
                p = C.c_void_p()
                _lib.check(lib.ssq_malloc(h, nbytes, C.byref(p)))
                self.mine[k] = p.value
                hb = (C.c_ubyte * 64)()
                _lib.check(lib.ssq_ipc_get_handle(h, p, hb))
                handles[64 * k: 64 * (k + 1)] = torch.frombuffer(bytes(hb), dtype=torch.uint8)
        except Exception as e:  # noqa: BLE001 -- the ranks agree on the outcome below
            failure = e
        mine = handles.to(self.ctx.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        if failure is None:
            try:
                for r in range(self.world):
                    hr = gathered[r].cpu().numpy().tobytes()
                    for k in range(4):
                        if r == self.rank:
                            self.peer[k][r] = self.mine[k]
                        else:
                            p = C.c_void_p()
                            _lib.check(lib.ssq_ipc_open(h, hr[64 * k: 64 * (k + 1)], C.byref(p)))
                            self.peer[k][r] = p.value
                            self._opened.append(p.value)
            except Exception as e:  # noqa: BLE001
                failure = e
        # every rank must take the same road: if CUDA IPC is not available to one of them, none uses it
        ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int64, device=self.ctx.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self._release()
            raise PeerExchangeUnavailable(f"CUDA IPC peer buffers could not be set up on every rank ({failure})")
        self.cap = cap
        self.rb_cap = rb_cap


def merge_peer(local, owner, exchange, group=None):
    """merge_alltoall with the exchange fused into the export: every rank's export kernel writes each owner's share
    directly into that owner's receive buffer (NVLink peer stores through `exchange`, a PeerExchange).  ShortSeq64
    counters; `owner` must have been created with hash_rot = log2(world).  Returns `owner`."""
    world, rank = exchange.world, exchange.rank
    t0 = _tick("start", time.perf_counter(), local.ctx) if _TIMING else 0.0
    parts = local.export_counts(world)                                  # tuples this rank holds for every owner
    mine = torch.cat([parts, torch.tensor([local.regions()], dtype=torch.int64, device=parts.device)])
    matrix = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(matrix, mine, group=group)                          # matrix[src] = tuples per dst, then src's region count
    mm = torch.stack(matrix).cpu().numpy()                              # host sync: every rank has entered this merge
    m, regions = mm[:, :world], mm[:, world] // world                   # sender regions per owner
    t0 = _tick("size matrix", t0, local.ctx)
    exchange.ensure(int(m.sum(axis=0).max()), int(regions.max()))
    import numpy as np
    before = m[:rank].sum(axis=0) if rank else np.zeros(world, dtype=np.int64)   # where my block starts in each owner's buffer
    table = np.empty((4, world), dtype=np.int64)
    for k, elem in enumerate((8, 1, 8)):
        table[k] = [exchange.peer[k][d] + elem * int(before[d]) for d in range(world)]
    rb_stride = exchange.rb_cap + 1
    table[3] = [exchange.peer[3][d] + 8 * rank * rb_stride for d in range(world)]
    dtable = torch.from_numpy(table).to(local.ctx.device)
    # rank r starts with owner r + 1: at any moment every owner receives from one sender
    local.export_to(world, dtable[:3].contiguous(), first_part=(rank + 1) % world)
    local.export_region_bases(world, dtable[3].contiguous())
    torch.cuda.synchronize(local.ctx.device)                            # my stores have landed ...
    dist.barrier(group)                                                 # ... and so have everyone else's
    t0 = _tick("export = exchange (peer stores)", t0, local.ctx)
    owner.merge_regions_raw(exchange.mine[0], exchange.mine[1], exchange.mine[2], int(m[:, rank].sum()), m[:, rank], regions,
                            exchange.mine[3], rb_stride)
    _tick("merge into owner table", t0, local.ctx)
    return owner


def global_size(owner, group=None):
    """Number of distinct keys over all ranks (sum of the disjoint owner tables)."""
    t = torch.tensor([len(owner)], dtype=torch.int64, device=owner.ctx.device)
    dist.all_reduce(t, group=group)
    return int(t.item())
