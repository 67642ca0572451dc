#!/usr/bin/env python3
"""bench.py -- Gbases/s of batched pack + dedup count (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config all|none|c1,c3,...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config.workload): BASELINE.json configs[1] ("C2") -- synthetic 32-nt reads -> ShortSeq64 pack +
dedup count, 1e9 reads and 1e8 distinct sequences PER GPU (weak scaling; the generator of SURVEY section 8d).
A step = one ShortSeqCounter construction over the whole resident batch: clear the table, the fused pack+count pass
over all reads (pack + level-1 scatter, region scatter, shared-memory region count; packed words and lengths are
written out), and for N > 1 the hash-partitioned exchange of the distinct keys (the export kernel stores each owner's
share into that owner's memory over NVLink; --nccl-exchange: NCCL all-to-all) and the merge into per-rank owner
tables; the step ends with the device->host read of the number of distinct keys.

  value        whole-job Gbases/s with the reads resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the host-buffer C-ABI call ssq_host_pack_count_lens: pinned host ASCII + one uint8
               length per read in; packed words AND the counter's result (keys, lengths, counts) back in pinned host
               memory inside the timed region (a bounded slice of the workload, size in e2e.reads_per_step)
  roofline     dominant kernel of the pass (most device time): algorithmic bytes per launch / its
               CUDA-event time (events recorded inside the library on the launching stream), against
               MEASURED_PEAKS.json hbm_gbs; "pass" = the whole fused pass (L+8+8W+1 per read + 8W+9 per unique),
               "kernels" = every kernel of the pass; traffic = DRAM bytes per launch from the committed ncu capture
  parity_check outside the timed region, at every N: sum of all owners' counts == reads processed, global number of
               distinct keys == distinct key ids of the generator (bincount on the device), and the multiplicities of
               65536 sampled reads per rank == the generator's bincount
  configs      the other BASELINE.json configs at reduced step counts: c1 (the Python drop-in beside the reference on
               1e6 x 22 nt), c2_u_sweep (U = 1e6 / 5e8), c3 (75-nt ShortSeq192; strong-scaled under torchrun), c4
               (ShortSeqVar 150/300/1000 mix pack + validate + decode round trip), c5 (Hamming pairs and ref-set)
  cpu_baseline / --impl reference: the unmodified reference (oracle/_ref, Cython) on the host cores

Only this file's cpu_baseline / --impl reference legs and the configs' parity samples touch oracle/; the measured path
never does.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED0001
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
ALL_CONFIGS = ("c1", "c2_u_sweep", "c3", "c4", "c5")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=float, default=1e9, help="reads per GPU")
    ap.add_argument("--uniques", type=float, default=1e8, help="distinct sequences per GPU shard's generator")
    ap.add_argument("--read-len", type=int, default=32)
    ap.add_argument("--e2e-reads", type=float, default=float(1 << 28), help="reads per GPU per e2e step")
    ap.add_argument("--cpu-reads", type=float, default=2e6, help="reads per CPU-baseline step (per process)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--config", default="all", help="extra configs: all, none, or a comma list of " + ",".join(ALL_CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every extra config by this factor (development)")
    ap.add_argument("--no-stream", action="store_true", help="N > 1: do not attach the counters (ssq_comm_attach): export + exchange after the pass")
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange the uniques with NCCL all-to-all instead of peer stores")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own ShortSeqCounter on the host cores
# ---------------------------------------------------------------------------------------------------
_SHARDS = None
_REF = None


def _count_shard(i):
    sq = _REF
    t0 = time.perf_counter()
    c = sq.ShortSeqCounter(_SHARDS[i])
    return time.perf_counter() - t0, len(c)


def make_cpu_lists(n_reads, read_len, dup_ratio, n_lists):
    """Bounded samples of the workload: n_lists lists of n_reads bytes objects from the shared generator,
    n_keys scaled so that reads/uniques stays at the workload's ratio."""
    from oracle import oracle as O
    n_keys = max(1, int(n_reads / dup_ratio))
    lists = []
    for s in range(n_lists):
        buf, off = O.synth_reads(SEED + 7919 * s, 0, n_reads, n_keys, read_len, read_len)
        raw = buf.tobytes()
        lists.append([raw[i * read_len:(i + 1) * read_len] for i in range(n_reads)])
    return lists


def reference_module():
    """The unmodified reference (oracle/_ref) when it was built here, else None."""
    from oracle import ref as R
    return R.load()


class OraclePortCounter:
    """Fallback CPU arm when oracle/_ref is absent: the C restatement (kind = "port")."""

    class _M:
        @staticmethod
        def ShortSeqCounter(reads):
            from oracle import oracle as O
            buf, off = O.concat(reads)
            w, l, _ = O.pack_batch(0 if len(reads[0]) <= 32 else 1, buf, off)
            return O.count(w, l, 1 if len(reads[0]) <= 32 else 3)[1]


def cpu_reference_run(args, steps, warmup):
    """Times ShortSeqCounter(list_of_bytes) of the reference: single process (its only mode) and one
    independent process per host core (per-shard counters, no merge -- an upper bound in the reference's
    favour).  Returns dict(value Gbases/s = the better of the two, ...)."""
    global _SHARDS, _REF
    import multiprocessing as mp
    ref = reference_module()
    kind = "reference"
    if ref is None:
        ref, kind = OraclePortCounter._M, "port"
    _REF = ref
    cores = os.cpu_count() or 1
    n = int(args.cpu_reads)
    L = args.read_len
    dup = args.reads / args.uniques
    _SHARDS = make_cpu_lists(n, L, dup, cores)
    single = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        c = ref.ShortSeqCounter(_SHARDS[0])
        dt = time.perf_counter() - t0
        if s >= warmup:
            single.append(dt)
        del c
    single_gb = n * L / statistics.mean(single) / 1e9
    multi_gb = 0.0
    multi_ms = None
    if cores > 1:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            walls = []
            for s in range(max(1, min(warmup, 1)) + max(2, min(steps, 3))):
                t0 = time.perf_counter()
                pool.map(_count_shard, range(cores), chunksize=1)
                walls.append(time.perf_counter() - t0)
            walls = walls[1:]
        multi_ms = statistics.mean(walls) * 1e3
        multi_gb = cores * n * L / statistics.mean(walls) / 1e9
    best_multi = multi_gb > single_gb
    _SHARDS = None
    return {
        "value": round(max(single_gb, multi_gb), 6), "unit": "Gbases/s", "cores": cores if best_multi else 1,
        "kind": kind,
        "sample": (f"{n} reads x {L} nt per process from the workload generator, n_keys scaled to keep "
                   f"reads/uniques = {dup:g}; single process {single_gb:.4f} Gbases/s; {cores} independent processes "
                   f"(per-shard counters, no merge) {multi_gb:.4f} Gbases/s"),
        "single_process_gbases_s": round(single_gb, 6), "all_cores_gbases_s": round(multi_gb, 6),
        "ms_per_step": round((multi_ms if best_multi else statistics.mean(single) * 1e3), 3),
    }


def cpu_c1_run(n=1_000_000, u=100_000, L=22):
    """Config C1 on the host, before CUDA is initialised: the reference's sq.pack loop and ShortSeqCounter(list) on
    1e6 x 22-nt reads (best of 3).  Returns the numbers and the reference's result as {sequence: count} (in order)."""
    from oracle import oracle as O
    ref = reference_module()
    if ref is None:
        return None
    buf, off = O.synth_reads(SEED, 0, n, u, L, L)
    raw = buf.tobytes()
    reads = [raw[i * L:(i + 1) * L] for i in range(n)]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        c = ref.ShortSeqCounter(reads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    items = [(str(k), v) for k, v in c.items()]
    sub = reads[:200_000]
    t0 = time.perf_counter()
    packed = [ref.pack(r) for r in sub]
    pack_us = (time.perf_counter() - t0) / len(sub) * 1e6
    t0 = time.perf_counter()
    _ = [str(s) for s in packed]
    str_us = (time.perf_counter() - t0) / len(sub) * 1e6
    t0 = time.perf_counter()
    _ = [a ^ b for a, b in zip(packed[:-1], packed[1:])]
    xor_us = (time.perf_counter() - t0) / (len(sub) - 1) * 1e6
    return {"counter_s": best, "items": items, "pack_us": pack_us, "str_us": str_us, "xor_us": xor_us, "n": n, "u": u, "L": L}


def config_of(args, world):
    n, u, L = int(args.reads), int(args.uniques), args.read_len
    return {
        "workload": f"{n:.3g} synthetic {L}-nt reads per GPU -> ShortSeq{'64' if L <= 32 else '192'} pack + dedup count "
                    f"({u:.3g} distinct sequences in the generator), BASELINE.json configs[1]",
        "reads_per_gpu": n, "distinct_sequences": u, "read_len": L,
        "parallelism": (f"dp{world}: reads sharded by index, local tables merged by a hash-partitioned exchange of the uniques "
                        f"({'torch.distributed all-to-all-v' if args.nccl_exchange else 'ssq_comm_attach + ssq_counter_merge_alltoall: the region count kernel stores every counted region into its owner over NVLink peer memory, device-side arrival flags' if not args.no_stream else 'ssq_counter_merge_alltoall: export kernel stores into the owners over NVLink peer memory, device-side arrival flags'})")
        if world > 1 else "single GPU",
        "l2_policy": "inputs (>= 40 GB per step) far exceed the 126 MB L2; no flush needed",
    }


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    r = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "Gbases/s pack+count", "value": r["value"], "unit": "Gbases/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_of(args, world),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], 0, None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        names = [n for b, n in self.REASONS.items() if self.reasons & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": names, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# helpers of our arm
# ---------------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full at the bench's full size)
    of each kernel of the pass from the committed capture, if any: {kernel name: bytes, "pass": bytes}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:  # noqa: BLE001
            pass
    return None


def _s64(x):
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= 1 << 63 else x


def _key_ids_at(x, n_keys, seed):
    """key_id(i) = mix64(seed + i) mod n_keys for the read numbers x (int64 tensor) -- the generator's key of read i
    (oracle/ssq_oracle.c:386-389, csrc/ssq_codec.cu synth_key) restated in torch int64 arithmetic (wrapping multiply,
    logical shifts by masking, unsigned modulo through the two 32-bit halves)."""
    x = x + _s64(seed)
    x = x ^ ((x >> 30) & ((1 << 34) - 1))
    x = x * _s64(0xBF58476D1CE4E5B9)
    x = x ^ ((x >> 27) & ((1 << 37) - 1))
    x = x * _s64(0x94D049BB133111EB)
    x = x ^ ((x >> 31) & ((1 << 33) - 1))
    hi, lo = (x >> 32) & 0xFFFFFFFF, x & 0xFFFFFFFF
    return (hi * ((1 << 32) % n_keys) + lo) % n_keys


def synth_key_ids(first, count, n_keys, seed, device):
    import torch
    return _key_ids_at(torch.arange(first, first + count, dtype=torch.int64, device=device), n_keys, seed)


def generator_bincount(first, n, n_keys, seed, device, world=1):
    """Multiplicity of every key id over this rank's reads [first, first + n), summed over all ranks -> int32 [n_keys]."""
    import torch
    import torch.distributed as dist
    bins = torch.zeros(n_keys, dtype=torch.int32, device=device)
    chunk = 1 << 26
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        ids = synth_key_ids(first + s, c, n_keys, seed, device)
        bins.index_add_(0, ids, torch.ones(c, dtype=torch.int32, device=device))
        del ids
    if world > 1:
        dist.all_reduce(bins)
    return bins


def parity_check(sq, table, words, lens, first, n, n_keys, seed, world, rank, klass, samples=1 << 16):
    """The correctness evidence printed with every bench line: `table` is the counter
    that holds this rank's share of the final result (the local table at N = 1, the owner table at N > 1); words / lens
    are this rank's packed reads [first, first + n) of the generator with n_keys keys."""
    import torch
    import torch.distributed as dist
    dev = words.device
    bins = generator_bincount(first, n, n_keys, seed, dev, world)
    expected_uniques = int((bins != 0).sum().item())
    keys, counts, _, _ = table.export(1)
    sums = torch.stack([counts.sum(), torch.tensor(len(table), dtype=torch.int64, device=dev)])
    del keys, counts
    if world > 1:
        dist.all_reduce(sums)
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + rank)
    idx = torch.randint(0, n, (samples,), generator=g).to(dev)
    s_words, s_lens = words[idx].contiguous(), lens[idx].contiguous()
    ids = _key_ids_at(idx + first, n_keys, seed)            # the generator evaluated at exactly those read numbers
    expect = bins[ids].to(torch.int64)
    if world > 1:
        gw = [torch.empty_like(s_words) for _ in range(world)]
        gl = [torch.empty_like(s_lens) for _ in range(world)]
        dist.all_gather(gw, s_words)
        dist.all_gather(gl, s_lens)
        allw, alll = torch.cat(gw), torch.cat(gl)
    else:
        allw, alll = s_words, s_lens
    got = table.lookup(sq.ShortSeqArray(table.ctx, klass, allw, alll))
    if world > 1:
        dist.all_reduce(got)
        got = got[rank * samples:(rank + 1) * samples]
    bad = int((got != expect).sum().item())
    if world > 1:
        t = torch.tensor([bad], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        bad = int(t.item())
    total_reads = world * n
    out = {"sum_counts": int(sums[0].item()), "expected_sum_counts": total_reads, "global_uniques": int(sums[1].item()),
           "expected_uniques": expected_uniques, "sampled": samples * world, "sampled_mismatches": bad,
           "how": "generator bincount(key_id) on the device vs export().counts.sum(), len(), and lookup() of sampled reads"}
    out["ok"] = bool(out["sum_counts"] == total_reads and out["global_uniques"] == expected_uniques and bad == 0)
    return out


class PassTimer:
    """Runs ssq_counter_pack_count steps on resident inputs and reports device times (CUDA events)."""

    def __init__(self, sq, klass, n, u, L, rank=0, first_read=None, expected_unique=None):
        import torch
        from shortseq_b200 import _lib
        self.sq, self.lib, self._lib = sq, _lib.lib(), _lib
        self.klass, self.n, self.u, self.L = klass, n, u, L
        self.W = 1 if klass == sq.CLASS_64 else 3
        self.first = rank * n if first_read is None else first_read
        self.batch = sq.synth_reads(n, u, L, L, seed=SEED, first_read=self.first)
        self.ctx = self.batch.ctx
        self.nbytes = int(self.batch.ascii.numel())
        self.words = self.ctx.empty((n,) if self.W == 1 else (n, 3), torch.int64)
        self.lens = self.ctx.empty((n,), torch.uint8)
        self.counter = sq.DeviceCounter(klass, expected_unique=expected_unique or u)
        self.phase_ms = []

    def step(self):
        import ctypes as C
        import torch
        from shortseq_b200._runtime import ptr
        lib, _lib = self.lib, self._lib
        _lib.check(lib.ssq_counter_clear(self.counter.handle))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ssq_counter_pack_count(self.counter.handle, ptr(self.batch.ascii), self.nbytes, ptr(self.batch.offsets),
                                              self.n, ptr(self.words), ptr(self.lens)))
        e1.record()
        d = [C.c_float(), C.c_float(), C.c_float()]
        _lib.check(lib.ssq_counter_last_pass_detail(self.counter.handle, C.byref(d[0]), C.byref(d[1]), C.byref(d[2])))
        self.phase_ms.append(tuple(x.value for x in d))
        return e0, e1

    def pass_bytes(self, uniques):
        """Algorithmic bytes of one fused pass: L + 8 + 8W + 1 per read + 8W + 9 per distinct key."""
        return self.n * (self.L + 8 + 8 * self.W + 1) + uniques * (8 * self.W + 9)


def timed(fn, steps, warmup):
    """Mean device milliseconds of fn() over `steps` calls after `warmup` (CUDA events on the current stream)."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def free_gpu():
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------
# the other BASELINE configs
# ---------------------------------------------------------------------------------------------------
def cfg_c2_u_sweep(sq, args, peak):
    """the C2 shape at U = 1e6 (table resident in L2, direct inserts) and U = 5e8 (2^30 slots)."""
    import torch
    out = []
    n = max(1 << 20, int(args.reads * args.scale))
    for u in (1e6, 5e8):
        u = max(1000, int(u * args.scale))
        free_gpu()
        try:
            pt = PassTimer(sq, sq.CLASS_64, n, u, args.read_len)
            ms = timed(lambda: (pt.step(), len(pt.counter)), 3, 2)
            uniq = len(pt.counter)
            rep = pt.ctx.sync()
            ph = [statistics.mean(p[i] for p in pt.phase_ms[-3:]) for i in range(3)]
            pc = parity_check(sq, pt.counter, pt.words, pt.lens, 0, n, u, SEED, 1, 0, sq.CLASS_64) if not args.no_parity else None
            out.append({"distinct_sequences": u, "reads": n, "table_slots": pt.counter.capacity(), "ms_per_step": round(ms, 3),
                        "gbases_s": round(n * args.read_len / ms / 1e6, 1), "uniques": uniq,
                        "pass_frac": round(pt.pass_bytes(uniq) / (ms * 1e-3) / 1e9 / peak, 4),
                        "kernel_ms": {"pack(+scatter/insert)": round(ph[0], 3), "region_scatter": round(ph[1], 3), "count": round(ph[2], 3)},
                        "status": rep.code, "parity_check": pc})
            del pt
        except Exception as e:  # noqa: BLE001 -- a config that cannot run is reported, not fatal
            out.append({"distinct_sequences": u, "error": repr(e)[:300]})
    free_gpu()
    return out


def cfg_c3(sq, args, peak, world, rank, state):
    """C3: 5e8 x 75-nt reads -> ShortSeq192 pack + count, U = N/16.  One GPU: the whole batch; under torchrun: strong
    scaling, rank r holds reads [r n/P, (r+1) n/P), the local tables merged into owner tables."""
    import torch
    import torch.distributed as dist
    from shortseq_b200.distributed import merge_alltoall
    total = max(1 << 22, int(5e8 * args.scale))
    n = total // world
    u = total // 16
    L = 75
    free_gpu()
    pt = PassTimer(sq, sq.CLASS_192, n, u, L, rank=rank, expected_unique=min(u, n))
    owner = None
    if world > 1:
        owner = sq.DeviceCounter(sq.CLASS_192, expected_unique=int(1.1 * u / world) + 1024, hash_rot=world.bit_length() - 1)

    xms = []

    def step():
        pt.step()
        if world > 1:
            pt._lib.check(pt.lib.ssq_counter_clear(owner.handle))
            if state["comm"] is not None:
                state["comm"].merge(pt.counter, owner)        # ShortSeq192: grouped ncclSend / ncclRecv inside the library
                xms.append((state["comm"].exchange_ms, state["comm"].merge_ms))
            else:
                merge_alltoall(pt.counter, owner=owner)
            return len(owner)
        return len(pt.counter)

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 3
    del pt.phase_ms[:]
    t0.record()
    for _ in range(K):
        step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / K
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=pt.ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rep = pt.ctx.sync()
    local_unique = len(pt.counter)
    ph = [statistics.mean(p[i] for p in pt.phase_ms) for i in range(3)]
    pass_ms = sum(ph)
    table = owner if world > 1 else pt.counter
    pc = None if args.no_parity else parity_check(sq, table, pt.words, pt.lens, rank * n, n, u, SEED, world, rank, sq.CLASS_192)
    res = {"workload": f"{total:.3g} synthetic 75-nt reads -> ShortSeq192 pack + dedup count, {u:.3g} distinct sequences"
                       + (f", strong-scaled over {world} GPUs (ssq_counter_merge_alltoall)" if world > 1 else ", 1 GPU"),
           "reads_total": total, "reads_per_gpu": n, "ms_per_step": round(ms, 3), "gbases_s": round(total * L / ms / 1e6, 1),
           "pass_ms_rank0": round(pass_ms, 3), "kernel_ms": {"pack+scatter": round(ph[0], 3), "count": round(ph[1] + ph[2], 3)},
           "pass_frac": round(pt.pass_bytes(local_unique) / (pass_ms * 1e-3) / 1e9 / peak, 4),
           "scaling": "strong", "status": rep.code, "parity_check": pc}
    if xms:
        res["exchange_ms"] = round(statistics.mean(x[0] for x in xms[-K:]), 3)
        res["merge_ms"] = round(statistics.mean(x[1] for x in xms[-K:]), 3)
    del pt, owner
    free_gpu()
    return res


def cfg_c4(sq, args, peak):
    """C4: ShortSeqVar pack + validate + decode round trip on a 150 / 300 / 1000-nt mix (50 / 30 / 20 %), all reads
    distinct, plus the 1 % bad-read variant (one N at a random position; the lowest bad read must be reported)."""
    import numpy as np
    import torch
    from shortseq_b200 import _lib
    from shortseq_b200._runtime import ptr
    from oracle import oracle as O
    free_gpu()
    n = max(1 << 16, int(5e7 * args.scale))
    ctx = sq.pack_batch([b"ACGT" * 30], klass=sq.CLASS_VAR).ctx
    dev = ctx.device
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0004)
    r = torch.rand(n, generator=g, device=dev)
    lens = torch.where(r < 0.5, 150, torch.where(r < 0.8, 300, 1000)).to(torch.int64)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_t = torch.empty(total, dtype=torch.uint8, device=dev)
    step_ = 1 << 30
    for s in range(0, total, step_):
        c = min(step_, total - s)
        ascii_t[s:s + c] = lut[torch.randint(0, 4, (c,), generator=g, device=dev, dtype=torch.uint8).long()]
    del r
    lib, h = _lib.lib(), ctx.bind()
    bound = lib.ssq_packvar_words_bound(total, n)
    words = ctx.empty((bound,), torch.int64)
    vlens = ctx.empty((n,), torch.int16)
    word_off = ctx.empty((n + 1,), torch.int64)
    out = ctx.empty((total,), torch.uint8)
    out_off = ctx.empty((n + 1,), torch.int64)

    def pack():
        _lib.check(lib.ssq_packvar(h, ptr(ascii_t), total, ptr(offsets), n, ptr(word_off), ptr(words), ptr(vlens)))

    def decode():
        _lib.check(lib.ssq_lens_to_offsets(h, ptr(vlens), 2, n, ptr(out_off)))
        _lib.check(lib.ssq_decodevar(h, ptr(words), ptr(word_off), ptr(vlens), n, ptr(out_off), ptr(out)))

    pack_ms = timed(pack, 3, 2)
    rep = ctx.sync()
    decode_ms = timed(decode, 3, 2)
    rep2 = ctx.sync()
    nwords = int(word_off[-1].item())
    round_trip = bool(torch.equal(out, ascii_t) and torch.equal(out_off, offsets))
    # a sample against the oracle: the first 20000 reads
    m = min(n, 20000)
    hb, ho = ascii_t[: int(offsets[m].item())].cpu().numpy(), offsets[: m + 1].cpu().numpy()
    ow, ol, owo = O.pack_batch(2, hb, ho)
    sample_ok = bool(np.array_equal(words[: int(word_off[m].item())].cpu().numpy().view(np.uint64), ow)
                     and np.array_equal(vlens[:m].cpu().numpy().astype(np.int64), ol.astype(np.int64)))
    pack_bytes = total + 8 * n + 8 * nwords + 8 * n + 2 * n          # ASCII + offset in; words + word offset + u16 length out
    decode_bytes = 8 * nwords + 2 * n + 8 * n + total + 8 * n        # words + length + word offset in; ASCII + offset out
    # 1 % bad reads: one 'N' somewhere in every 100th read; the report must name the lowest one
    bad_reads = torch.arange(37, n, 100, device=dev)
    pos = offsets[bad_reads] + (torch.rand(bad_reads.numel(), generator=g, device=dev) * lens[bad_reads]).long()
    saved = ascii_t[pos].clone()
    ascii_t[pos] = ord("N")
    bad_ms = timed(pack, 2, 1)
    rep3 = ctx.sync()
    ascii_t[pos] = saved
    res = {"workload": f"{n:.3g} ShortSeqVar reads, 150/300/1000 nt at 50/30/20 % (mean {total / n:.0f} nt, all distinct): pack + validate + decode",
           "reads": n, "bases": total, "pack_ms": round(pack_ms, 3), "decode_ms": round(decode_ms, 3),
           "round_trip_gbases_s": round(total / (pack_ms + decode_ms) / 1e6, 1),
           "pack_frac": round(pack_bytes / (pack_ms * 1e-3) / 1e9 / peak, 4), "decode_frac": round(decode_bytes / (decode_ms * 1e-3) / 1e9 / peak, 4),
           "round_trip_frac": round((pack_bytes + decode_bytes) / ((pack_ms + decode_ms) * 1e-3) / 1e9 / peak, 4),
           "algorithmic_bytes": {"pack": pack_bytes, "decode": decode_bytes},
           "bad_read_variant": {"pack_ms": round(bad_ms, 3), "reported_status": rep3.code, "reported_first_bad_read": int(rep3.first_bad_read),
                                "expected_first_bad_read": 37},
           "parity": {"round_trip_identical": round_trip, "oracle_sample_reads": m, "oracle_sample_identical": sample_ok,
                      "status": [rep.code, rep2.code],
                      "ok": bool(round_trip and sample_ok and rep.code == 0 and rep2.code == 0 and rep3.code == _lib.ERR_BAD_BASE
                                 and int(rep3.first_bad_read) == 37)}}
    del ascii_t, words, out, offsets, out_off, word_off, vlens, lens
    free_gpu()
    return res


def cfg_c5(sq, args, peak):
    """C5: batched Hamming distance of packed 12-nt UMIs (W = 1) and 96-nt reads (W = 3): element-wise pairs, and every
    query against a reference set of 64 / 1024 sequences (min, argmin, number within distance 1)."""
    import numpy as np
    import torch
    from oracle import oracle as O
    out = {}
    n = max(1 << 20, int(1e8 * args.scale))
    for name, L, klass, W in (("umi12", 12, sq.CLASS_64, 1), ("read96", 96, sq.CLASS_192, 3)):
        free_gpu()
        a = sq.pack_batch(sq.synth_reads(n, n, L, L, seed=SEED), klass=klass)
        b = sq.pack_batch(sq.synth_reads(n, n, L, L, seed=SEED + 99), klass=klass)
        ms = timed(lambda: sq.hamming_batch(a, b), 3, 2)
        d = sq.hamming_batch(a, b)
        m = 100_000
        aw, al, _ = sq.ShortSeqArray(a.ctx, klass, a.words[:m].contiguous(), a.lens[:m].contiguous()).to_host()
        bw, bl, _ = sq.ShortSeqArray(a.ctx, klass, b.words[:m].contiguous(), b.lens[:m].contiguous()).to_host()
        ok = bool(np.array_equal(d[:m].cpu().numpy().astype(np.int32), O.hamming_batch(aw, bw, al, bl, W)))
        pair_bytes = n * (16 * W + 1 + 1)          # both keys + one length in, one byte out
        res = {"pairs": n, "pairs_ms": round(ms, 3), "gpairs_s": round(n / ms / 1e6, 2),
               "pairs_frac_hbm": round(pair_bytes / (ms * 1e-3) / 1e9 / peak, 4), "bytes_per_pair": 16 * W + 2, "parity_sample_ok": ok}
        for R in (64, 1024):
            refs = sq.pack_batch(sq.synth_reads(R, R, L, L, seed=SEED + 7), klass=klass)
            ms = timed(lambda: sq.hamming_refset(a, refs, thresh=1), 2, 1)
            md, am, within = sq.hamming_refset(a, refs, thresh=1)
            # parity: the first 2000 queries against the oracle's distances to every reference
            q = 2000
            rw, rl, _ = refs.to_host()
            best = np.full(q, 255, np.int32)
            for j in range(R):
                dj = O.hamming_batch(aw[:q], np.repeat(rw[j:j + 1], q, axis=0), al[:q], np.repeat(rl[j:j + 1], q), W)
                best = np.minimum(best, dj)
            rok = bool(np.array_equal(md[:q].cpu().numpy().astype(np.int32), best))
            cmp_s = n * R / (ms * 1e-3)
            res[f"refset_{R}"] = {"ms": round(ms, 3), "gcomparisons_s": round(cmp_s / 1e9, 1),
                                  "frac_hbm": round(n * (8 * W + 1 + 9) / (ms * 1e-3) / 1e9 / peak, 4),
                                  "bound": "ALU (xor + collapse + popc per 32-base block, R per query); HBM fraction shown for reference",
                                  "parity_sample_ok": rok}
            del refs
        out[name] = res
        del a, b, d
    free_gpu()
    # the consumer of the UMI Hamming kernel (row N3): UMI-tools' directional clustering of the distinct 12-nt UMIs of many
    # groups (one group per mapping position): all pairs inside a group, label propagation until stable, one CTA per group
    try:
        groups, per = max(1024, int(2e5 * args.scale)), 32
        nu = groups * per
        umis = sq.pack_batch(sq.synth_reads(nu, nu, 12, 12, seed=SEED + 123), klass=sq.CLASS_64)
        cnt = torch.randint(1, 50, (nu,), dtype=torch.int64, device=umis.words.device)
        goff = torch.arange(0, nu + 1, per, dtype=torch.int64, device=umis.words.device)
        ms = timed(lambda: sq.umi_collapse(umis, cnt, goff, 1, "directional"), 3, 1)
        rep, ccounts, ncl = sq.umi_collapse(umis, cnt, goff, 1, "directional")
        out["umi_collapse"] = {"groups": groups, "umis_per_group": per, "threshold": 1, "method": "directional", "ms": round(ms, 3),
                               "gpairs_per_sweep_s": round(groups * per * per / (ms * 1e-3) / 1e9, 1),
                               "mumis_s": round(nu / ms / 1e3, 1), "clusters": int(ncl.sum().item()),
                               "counts_conserved": bool(int(ccounts.sum().item()) == int(cnt.sum().item()))}
        del umis, cnt, goff, rep, ccounts, ncl
    except Exception as e:  # noqa: BLE001 -- reported, not fatal
        out["umi_collapse"] = {"error": repr(e)[:300]}
    free_gpu()
    return out


def cfg_c1(sq, c1_cpu):
    """C1: the Python drop-in beside the reference on the reference's own shape (1e6 x 22-nt reads, 1e5 distinct):
    ShortSeqCounter(list_of_bytes) wall time (best of 3) and the per-call costs of sq.pack / str / ^."""
    from oracle import oracle as O
    if c1_cpu is None:
        return {"error": "reference not built (oracle/_ref missing)"}
    n, u, L = c1_cpu["n"], c1_cpu["u"], c1_cpu["L"]
    buf, off = O.synth_reads(SEED, 0, n, u, L, L)
    raw = buf.tobytes()
    reads = [raw[i * L:(i + 1) * L] for i in range(n)]
    best = None
    for _ in range(4):
        t0 = time.perf_counter()
        c = sq.ShortSeqCounter(reads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    items = [(str(k), v) for k, v in c.items()]
    same = items == c1_cpu["items"]
    sub = reads[:200_000]
    t0 = time.perf_counter()
    packed = [sq.pack(r) for r in sub]
    pack_us = (time.perf_counter() - t0) / len(sub) * 1e6
    t0 = time.perf_counter()
    _ = [str(s) for s in packed]
    str_us = (time.perf_counter() - t0) / len(sub) * 1e6
    t0 = time.perf_counter()
    _ = [a ^ b for a, b in zip(packed[:-1], packed[1:])]
    xor_us = (time.perf_counter() - t0) / (len(sub) - 1) * 1e6
    return {"workload": f"sq.pack + ShortSeqCounter(list) on {n} synthetic {L}-nt reads, {u} distinct (launch-latency bound: wall time, not a roofline)",
            "ours_counter_s": round(best, 4), "reference_counter_s": round(c1_cpu["counter_s"], 4),
            "ours_gbases_s": round(n * L / best / 1e9, 4), "reference_gbases_s": round(n * L / c1_cpu["counter_s"] / 1e9, 4),
            "per_call_us": {"ours": {"pack": round(pack_us, 3), "str": round(str_us, 3), "xor": round(xor_us, 3)},
                            "reference": {"pack": round(c1_cpu["pack_us"], 3), "str": round(c1_cpu["str_us"], 3), "xor": round(c1_cpu["xor_us"], 3)}},
            "parity": {"same_items_same_order": bool(same), "distinct": len(items)}}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, emit):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wanted = set(ALL_CONFIGS) if args.config == "all" else set(x for x in args.config.split(",") if x and x != "none")

    cpu = c1_cpu = None
    if rank == 0 and world == 1:                            # before CUDA is initialised (the CPU legs fork)
        if not args.no_cpu_baseline:
            cpu = cpu_reference_run(args, steps=3, warmup=1)
        if "c1" in wanted:
            c1_cpu = cpu_c1_run()

    import ctypes as C
    import torch
    import torch.distributed as dist
    import shortseq_b200 as sq
    from shortseq_b200 import _lib
    from shortseq_b200.distributed import Comm, merge_alltoall

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.lib()
    n, u, L = int(args.reads), int(args.uniques), args.read_len
    klass = sq.CLASS_64 if L <= 32 else sq.CLASS_192
    W = 1 if klass == sq.CLASS_64 else 3
    K, WU = args.steps, max(3, args.warmup)

    # resident inputs: rank r holds reads [r*n, (r+1)*n) of the global generator
    pt = PassTimer(sq, klass, n, u, L, rank=rank)
    ctx, local = pt.ctx, pt.counter
    owner = comm = None
    if world > 1:
        # every rank draws from the same u keys, and owners split the key space evenly by hash
        owner = sq.DeviceCounter(klass, expected_unique=int(1.1 * u / world) + 1024, hash_rot=world.bit_length() - 1)
        if not args.nccl_exchange:
            comm = Comm(ctx)            # the library's own communicator: NCCL for the sizes, NVLink peer stores for the payload
            if not args.no_stream:
                try:
                    comm.attach(local, owner)   # streamed exchange: the count kernel sends every table region to its owner as it goes
                except Exception as e:  # noqa: BLE001 -- the decision is collective (every rank sees the same matrix): all ranks land here together
                    print(f"[bench] streamed exchange unavailable, merging unstreamed: {e!r}", file=sys.stderr)
    state = {"comm": comm}
    h = ctx.bind()
    uniques_seen = [0]
    xms = []

    def step():
        ev = pt.step()
        if world > 1:
            _lib.check(lib.ssq_counter_clear(owner.handle))
            if comm is not None:
                comm.merge(local, owner)                  # ssq_counter_merge_alltoall
                xms.append((comm.exchange_ms, comm.merge_ms))
            else:
                merge_alltoall(local, owner=owner)        # export, then torch.distributed all-to-all-v
            uniques_seen[0] = len(owner)
        else:
            uniques_seen[0] = len(local)          # device->host read of the step's result
        return ev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(WU):
        step()
    rep = ctx.sync()
    assert rep.code == 0, f"device reported status {rep.code} at read {rep.first_bad_read}"
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.ssq_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    del pt.phase_ms[:]
    del xms[:]
    evs = [step() for _ in range(K)]
    t1.record()
    barrier()
    launches = lib.ssq_launch_count() - launches0
    clocks = sampler.stop()
    total_ms = t0.elapsed_time(t1)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    local_unique = len(local)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        tl = torch.tensor([launches], dtype=torch.int64, device=ctx.device)
        dist.all_reduce(tl)
        launches = int(tl.item())
    ms_per_step = total_ms / K
    value = world * n * L / (ms_per_step * 1e-3) / 1e9

    # parity of the step just timed (outside the timed region)
    pc = None
    if not args.no_parity:
        pc = parity_check(sq, owner if world > 1 else local, pt.words, pt.lens, rank * n, n, u, SEED, world, rank, klass)

    # roofline (this rank).  Algorithmic bytes are SURVEY 8d's: pack = ASCII + offset in, words + len out;
    # count = the packed keys in, one (key, len, count) tuple per distinct key out.
    peak, peak_src = hbm_peak()
    pack_bytes = n * (L + 8 + 8 * W + 1)
    count_bytes = n * (8 * W + 1) + local_unique * (8 * W + 9)
    alg_bytes = pt.pass_bytes(local_unique)       # the fused pass: packed keys never re-read algorithmically
    k_ms = statistics.mean(kernel_ms)
    p_ms = [statistics.mean(p[i] for p in pt.phase_ms) for i in range(3)]
    traffic = ncu_traffic() or {}
    deferred = p_ms[1] + p_ms[2] > 0
    regions = p_ms[1] > 0

    def kern(name, what, ms, bytes_):
        d = {"kernel": name, "what": what, "ms_per_launch": round(ms, 3), "launches_per_step": 1,
             "algorithmic_bytes_per_launch": bytes_, "achieved_gbs": round(bytes_ / (ms * 1e-3) / 1e9, 1) if bytes_ else None,
             "traffic": traffic.get(name.split("<")[0])}
        d["frac"] = round(d["achieved_gbs"] / peak, 4) if bytes_ else None
        return d

    pack_name = "ssq::pack32_kernel<{mode},256>" if (klass == sq.CLASS_64 and L == 32) else f"ssq::pack_fixed_kernel<{W // 3},{{mode}}>"
    if deferred:
        kernels = [kern(pack_name.format(mode=2), "pack + validate + scatter keys to 256 hash partitions", p_ms[0], pack_bytes)]
        if regions:
            kernels.append(kern("ssq::region_scatter_kernel", "route keys to their 4096-slot table region (second 256-way scatter); "
                                "shares the count phase's algorithmic bytes with the region count", p_ms[1], 0))
            kernels.append(kern("ssq::count_regions2_kernel<384,%s>" % ("true" if (comm is not None and comm.streams) else "false"),
                                "count each region's keys in shared memory, write the table back"
                                + (" and store the region's (key, count) pairs into the owner rank's receive buffer (streamed exchange)"
                                   if (comm is not None and comm.streams) else ""), p_ms[2], count_bytes))
            kernels[-1]["achieved_gbs"] = round(count_bytes / ((p_ms[1] + p_ms[2]) * 1e-3) / 1e9, 1)   # over both count-phase kernels
            kernels[-1]["frac"] = round(kernels[-1]["achieved_gbs"] / peak, 4)
        else:
            kernels.append(kern("ssq::count_parts_kernel", "partition-ordered table insertion (L2-resident table ranges)", p_ms[2], count_bytes))
    else:
        kernels = [kern(pack_name.format(mode=1), "pack + validate + insert", p_ms[0], alg_bytes)]
    dom = max(kernels, key=lambda kk: kk["ms_per_launch"])
    # top level = the dominant kernel (most device time per step); "pass" = all kernels of one ssq_counter_pack_count
    roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "frac_of_8tbs_spec": round(dom["achieved_gbs"] / 8000.0, 4),
                "traffic": dom["traffic"], "kernel": dom["kernel"], "kernel_ms_per_launch": dom["ms_per_launch"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "peak_source": peak_src,
                "pass": {"kernels": " + ".join(kk["kernel"] for kk in kernels), "ms": round(k_ms, 3), "algorithmic_bytes": alg_bytes,
                         "achieved": round(alg_bytes / (k_ms * 1e-3) / 1e9, 1), "frac": round(alg_bytes / (k_ms * 1e-3) / 1e9 / peak, 4),
                         "traffic": traffic.get("pass"),
                         "traffic_over_algorithmic": round(traffic["pass"] / alg_bytes, 3) if traffic.get("pass") else None},
                "kernels": kernels}

    # the headline's buffers are no longer needed
    streamed = bool(comm is not None and comm.last_streamed)
    if comm is not None and comm.streams:
        comm.attach(None, None)         # collective: frees the streamed exchange's receive buffers
    table_slots = local.capacity()
    del pt, local, owner
    free_gpu()

    # e2e: host buffers through the C ABI, results back in host memory
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(sq, args, world, rank, state, barrier, h)

    extra = {}
    if world == 1:
        if "c1" in wanted:
            extra["c1"] = cfg_c1(sq, c1_cpu)
        if "c2_u_sweep" in wanted:
            extra["c2_u_sweep"] = cfg_c2_u_sweep(sq, args, peak)
        if "c4" in wanted:
            extra["c4"] = cfg_c4(sq, args, peak)
        if "c5" in wanted:
            extra["c5"] = cfg_c5(sq, args, peak)
    if "c3" in wanted:
        extra["c3"] = cfg_c3(sq, args, peak, world, rank, state)

    if rank == 0:
        line = {
            "metric": "Gbases/s pack+count", "value": round(value, 2), "unit": "Gbases/s", "n_gpus": world, "steps": K,
            "warmup": WU, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_of(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "uniques": int(uniques_seen[0]), "table_slots": table_slots, "parity_check": pc, "configs": extra,
        }
        if world > 1 and xms:
            line["exchange_ms"] = round(statistics.mean(x[0] for x in xms), 3)      # rank 0: send side (export kernel = peer stores)
            line["merge_ms"] = round(statistics.mean(x[1] for x in xms), 3)         # rank 0: wait for the senders + owner-side count
            line["exchange"] = ("streamed: count_regions2_kernel stores every counted region into its owner's memory over NVLink (ssq_comm_attach); "
                                "the merge publishes arrival flags and adds the blocks up" if streamed else
                                "peer stores over NVLink (ssq_counter_merge_alltoall)" if comm.peer_stores else "grouped ncclSend/ncclRecv (ssq_counter_merge_alltoall)")
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


def run_e2e(sq, args, world, rank, state, barrier, h):
    """The same metric through the reference-facing C call with HOST buffers: pinned ASCII + one uint8 length per read
    in; packed words out; then the counter's result -- (key, length, count) of every distinct sequence -- is exported
    and copied back to pinned host memory, all inside the timed region."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from shortseq_b200 import _lib
    from shortseq_b200.distributed import merge_alltoall
    lib = _lib.lib()
    n, u, L = int(args.reads), int(args.uniques), args.read_len
    klass = sq.CLASS_64 if L <= 32 else sq.CLASS_192
    W = 1 if klass == sq.CLASS_64 else 3
    K = args.steps
    ne = int(min(n, args.e2e_reads))
    ue = max(1, int(ne / (n / u)))
    chunk = 1 << 26
    h_ascii = torch.empty(ne * L, dtype=torch.uint8).pin_memory()
    for s in range(0, ne, chunk):                              # fill the pinned buffer slice by slice
        c = min(chunk, ne - s)
        eb = sq.synth_reads(c, ue, L, L, seed=SEED, first_read=rank * ne + s)
        h_ascii[s * L:(s + c) * L].copy_(eb.ascii[: c * L])
        del eb
    ctx = sq.pack_batch([b"ACGT"], klass=sq.CLASS_64).ctx
    h_lens = torch.full((ne,), L, dtype=torch.uint8).pin_memory()      # one length per read, as a list of bytes carries
    h_words = torch.empty((ne,) if W == 1 else (ne, 3), dtype=torch.int64).pin_memory()
    cap_u = int(1.05 * ue) + 1024
    r_words = torch.empty((cap_u,) if W == 1 else (cap_u, 3), dtype=torch.int64).pin_memory()
    r_lens = torch.empty((cap_u,), dtype=torch.uint8).pin_memory()
    r_counts = torch.empty((cap_u,), dtype=torch.int64).pin_memory()
    ectr = sq.DeviceCounter(klass, expected_unique=ue)
    eowner = sq.DeviceCounter(klass, expected_unique=int(1.1 * ue / world) + 1024, hash_rot=world.bit_length() - 1) if world > 1 else None
    rep = _lib.Report()
    d2h = [0]

    def estep():
        _lib.check(lib.ssq_counter_clear(ectr.handle))
        _lib.check(lib.ssq_host_pack_count_lens(h, ectr.handle, h_ascii.data_ptr(), h_lens.data_ptr(), ne, h_words.data_ptr(),
                                                1 << 22, C.byref(rep)))
        assert rep.code == 0
        table = ectr
        if world > 1:
            _lib.check(lib.ssq_counter_clear(eowner.handle))
            if state["comm"] is not None:
                state["comm"].merge(ectr, eowner)
            else:
                merge_alltoall(ectr, owner=eowner)
            table = eowner
        keys, counts, _, _ = table.export(1)                    # the step's result comes back to the host
        m = len(keys)
        r_words[:m].copy_(keys.words, non_blocking=True)
        r_lens[:m].copy_(keys.lens, non_blocking=True)
        r_counts[:m].copy_(counts, non_blocking=True)
        torch.cuda.synchronize()
        d2h[0] = ne * 8 * W + m * (8 * W + 1 + 8)
        return m

    for _ in range(2):
        estep()
    barrier()
    w0 = time.perf_counter()
    for _ in range(K):
        estep()
    barrier()
    e_ms = (time.perf_counter() - w0) * 1e3 / K
    if world > 1:
        t = torch.tensor([e_ms], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    h2d = ne * L + ne
    res = {"value": round(world * ne * L / (e_ms * 1e-3) / 1e9, 3), "unit": "Gbases/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h[0],
           "reads_per_step": ne, "ms_per_step": round(e_ms, 3),
           "host_link_gbs_per_gpu": round((h2d + d2h[0]) / (e_ms * 1e-3) / 1e9, 1),
           "call": "ssq_host_pack_count_lens (pinned host ASCII + one uint8 length per read in, packed words out, 4M-read "
                   "chunks, H2D / kernel / D2H overlapped on three streams), then export of (keys, lengths, counts) to pinned host memory",
           "sample": f"{ne} reads per GPU per step: a slice of the {int(args.reads):.3g}-read workload from the same generator with the "
                     "reads-per-distinct-key ratio kept (33 B of pinned host memory per read; the rate is PCIe-bound and does not depend on the slice)"}
    del ectr, eowner, h_ascii, h_words
    free_gpu()
    return res


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one): everything
    # written to file descriptor 1 while the bench runs goes to stderr, and only the final line to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")

    def emit(line):
        out.write(json.dumps(line) + "\n")
        out.flush()

    if args.impl == "reference":
        run_reference(args, emit)
    else:
        run_ours(args, emit)


if __name__ == "__main__":
    main()
