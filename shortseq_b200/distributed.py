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


def global_size(owner, group=None):
    """Number of distinct keys over all ranks (sum of the disjoint owner tables)."""
    t = torch.tensor([len(owner)], dtype=torch.int64, device=owner.ctx.device)
    dist.all_reduce(t, group=group)
    return int(t.item())
