#!/usr/bin/env python3
"""bench.py -- Gbases/s of batched pack + dedup count (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- synthetic 32-nt reads -> ShortSeq64 pack + dedup
count, 1e9 reads and 1e8 distinct sequences PER GPU (weak scaling; the generator of SURVEY section 8d).
A step = one ShortSeqCounter construction over the whole resident batch: clear the table, the fused
pack+count pass over all reads (pack + level-1 scatter, region scatter, shared-memory region count; packed
words and lengths are written out), and for N > 1 the hash-partitioned exchange of the distinct keys (the export
kernel stores each owner's share into that owner's memory over NVLink; --nccl-exchange: NCCL all-to-all) and the
merge into per-rank owner tables; the step ends with the device->host read of the number of distinct keys.

  value     whole-job Gbases/s with the reads resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the host-buffer C-ABI call ssq_host_pack_count_lens: pinned host ASCII + one
            uint8 length per read in, packed words out, host<->device copies inside the timed region (a
            bounded slice of the workload, size in e2e.reads_per_step)
  roofline  dominant kernel of the pass (most device time): algorithmic bytes per launch / its
            CUDA-event time (events recorded inside the library on the launching stream), against
            MEASURED_PEAKS.json hbm_gbs; "pass" = the whole fused pass (L+8+8W+1 per read + 8W+9 per unique),
            "kernels" = every kernel of the pass; traffic = DRAM bytes per launch from the committed ncu capture
  cpu_baseline / --impl reference: the unmodified reference (oracle/_ref, Cython) on the host cores

Only this file's cpu_baseline / --impl reference legs touch oracle/; the measured path never does.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=float, default=1e9, help="reads per GPU")
    ap.add_argument("--uniques", type=float, default=1e8, help="distinct sequences per GPU shard's generator")
    ap.add_argument("--read-len", type=int, default=32)
    ap.add_argument("--e2e-reads", type=float, default=float(1 << 27), help="reads per GPU per e2e step")
    ap.add_argument("--cpu-reads", type=float, default=2e6, help="reads per CPU-baseline step (per process)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange the uniques with NCCL all-to-all instead of peer stores")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own ShortSeqCounter on the host cores
# ---------------------------------------------------------------------------------------------------
_SHARDS = None


def _count_shard(i):
    sq = _REF
    t0 = time.perf_counter()
    c = sq.ShortSeqCounter(_SHARDS[i])
    return time.perf_counter() - t0, len(c)


_REF = None


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
    # single process
    single = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        c = ref.ShortSeqCounter(_SHARDS[0])
        dt = time.perf_counter() - t0
        if s >= warmup:
            single.append(dt)
        del c
    single_gb = n * L / statistics.mean(single) / 1e9
    # all cores: one forked process per core, each on its own shard
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
    return {
        "value": round(max(single_gb, multi_gb), 6), "unit": "Gbases/s", "cores": cores if best_multi else 1,
        "kind": kind,
        "sample": (f"{n} reads x {L} nt per process from the workload generator, n_keys scaled to keep "
                   f"reads/uniques = {dup:g}; single process {single_gb:.4f} Gbases/s; {cores} independent processes "
                   f"(per-shard counters, no merge) {multi_gb:.4f} Gbases/s"),
        "single_process_gbases_s": round(single_gb, 6), "all_cores_gbases_s": round(multi_gb, 6),
        "ms_per_step": round((multi_ms if best_multi else statistics.mean(single) * 1e3), 3),
    }


def config_of(args, world):
    n, u, L = int(args.reads), int(args.uniques), args.read_len
    return {
        "workload": f"{n:.3g} synthetic {L}-nt reads per GPU -> ShortSeq{'64' if L <= 32 else '192'} pack + dedup count "
                    f"({u:.3g} distinct sequences in the generator), BASELINE.json configs[1]",
        "reads_per_gpu": n, "distinct_sequences": u, "read_len": L,
        "parallelism": (f"dp{world}: reads sharded by index, local tables merged by a hash-partitioned exchange of the uniques "
                        f"({'NCCL all-to-all' if args.nccl_exchange else 'export kernel stores into the owners over NVLink peer memory'})")
        if world > 1 else "single GPU",
        "l2_policy": "inputs (>= 40 GB per step) far exceed the 126 MB L2; no flush needed",
    }


def run_reference(args):
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
    print(json.dumps(line), flush=True)


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
# our arm
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


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(args, steps=3, warmup=1)     # before CUDA is initialised (it forks)

    import ctypes as C
    import torch
    import torch.distributed as dist
    import shortseq_b200 as sq
    from shortseq_b200 import _lib
    from shortseq_b200._runtime import ptr
    from shortseq_b200.distributed import PeerExchange, PeerExchangeUnavailable, merge_alltoall, merge_peer

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.lib()
    n, u, L = int(args.reads), int(args.uniques), args.read_len
    klass = sq.CLASS_64 if L <= 32 else sq.CLASS_192
    W = 1 if klass == sq.CLASS_64 else 3
    K, WU = args.steps, max(3, args.warmup)

    # resident inputs: rank r holds reads [r*n, (r+1)*n) of the global generator
    batch = sq.synth_reads(n, u, L, L, seed=SEED, first_read=rank * n)
    ctx = batch.ctx
    nbytes = int(batch.ascii.numel())
    words = ctx.empty((n,) if W == 1 else (n, 3), torch.int64)
    lens = ctx.empty((n,), torch.uint8)
    local = sq.DeviceCounter(klass, expected_unique=u)
    owner = exchange = None
    if world > 1:
        # every rank draws from the same u keys, and owners split the key space evenly by hash
        owner = sq.DeviceCounter(klass, expected_unique=int(1.1 * u / world) + 1024, hash_rot=world.bit_length() - 1)
        if klass == sq.CLASS_64 and not args.nccl_exchange:
            exchange = PeerExchange(ctx)
    state = {"exchange": exchange}
    h = ctx.bind()
    kernel_ms, uniques_seen, phase_ms = [], [0], []

    def step():
        _lib.check(lib.ssq_counter_clear(local.handle))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ssq_counter_pack_count(local.handle, ptr(batch.ascii), nbytes, ptr(batch.offsets), n, ptr(words), ptr(lens)))
        e1.record()
        d = [C.c_float(), C.c_float(), C.c_float()]
        _lib.check(lib.ssq_counter_last_pass_detail(local.handle, C.byref(d[0]), C.byref(d[1]), C.byref(d[2])))   # CUDA events inside the library
        phase_ms.append(tuple(x.value for x in d))
        if world > 1:
            _lib.check(lib.ssq_counter_clear(owner.handle))
            if state["exchange"] is not None:
                try:
                    merge_peer(local, owner, state["exchange"])   # export kernel stores straight into the owners' memory (NVLink)
                except PeerExchangeUnavailable as e:              # raised on every rank alike: switch to NCCL for good
                    if rank == 0:
                        print(f"bench: {e}; using the NCCL all-to-all exchange", file=sys.stderr, flush=True)
                    state["exchange"] = None
                    _lib.check(lib.ssq_counter_clear(owner.handle))
                    merge_alltoall(local, owner=owner)
            else:
                merge_alltoall(local, owner=owner)        # export, then NCCL all-to-all
            uniques_seen[0] = len(owner)
        else:
            uniques_seen[0] = len(local)          # device->host read of the step's result
        return e0, e1

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
    del phase_ms[:]
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

    # roofline (this rank).  Algorithmic bytes are SURVEY 8d's: pack = ASCII + offset in, words + len out;
    # count = the packed keys in, one (key, len, count) tuple per distinct key out.
    peak, peak_src = hbm_peak()
    pack_bytes = n * (L + 8 + 8 * W + 1)
    count_bytes = n * (8 * W + 1) + local_unique * (8 * W + 9)
    alg_bytes = n * (L + 8 + 8 * W + 1) + local_unique * (8 * W + 9)       # the fused pass: packed keys never re-read algorithmically
    k_ms = statistics.mean(kernel_ms)
    p_ms = [statistics.mean(p[i] for p in phase_ms) for i in range(3)]
    traffic = ncu_traffic() or {}
    deferred = p_ms[1] + p_ms[2] > 0
    regions = p_ms[1] > 0

    def kern(name, what, ms, bytes_):
        d = {"kernel": name, "what": what, "ms_per_launch": round(ms, 3), "launches_per_step": 1,
             "algorithmic_bytes_per_launch": bytes_, "achieved_gbs": round(bytes_ / (ms * 1e-3) / 1e9, 1) if bytes_ else None,
             "traffic": traffic.get(name.split("<")[0])}
        d["frac"] = round(d["achieved_gbs"] / peak, 4) if bytes_ else None
        return d

    if deferred:
        kernels = [kern(f"ssq::pack_fixed_kernel<{W // 3},2>", "pack + validate + scatter keys to 256 hash partitions", p_ms[0], pack_bytes)]
        if regions:
            kernels.append(kern("ssq::region_scatter_kernel", "route keys to their 4096-slot table region (second 256-way scatter); "
                                "shares the count phase's algorithmic bytes with count_regions_kernel", p_ms[1], 0))
            kernels.append(kern("ssq::count_regions_kernel<384>", "count each region's keys in shared memory, write the table back",
                                p_ms[2], count_bytes))
            kernels[-1]["achieved_gbs"] = round(count_bytes / ((p_ms[1] + p_ms[2]) * 1e-3) / 1e9, 1)   # over both count-phase kernels
            kernels[-1]["frac"] = round(kernels[-1]["achieved_gbs"] / peak, 4)
        else:
            kernels.append(kern("ssq::count_parts_kernel", "partition-ordered table insertion (L2-resident table ranges)", p_ms[2], count_bytes))
    else:
        kernels = [kern(f"ssq::pack_fixed_kernel<{W // 3},1>", "pack + validate + insert", p_ms[0], alg_bytes)]
    dom = max(kernels, key=lambda kk: kk["ms_per_launch"])
    # top level = the dominant kernel (most device time per step); "pass" = all kernels of one ssq_counter_pack_count
    roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "frac_of_8tbs_spec": round(dom["achieved_gbs"] / 8000.0, 4),
                "traffic": dom["traffic"], "kernel": dom["kernel"], "kernel_ms_per_launch": dom["ms_per_launch"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "peak_source": peak_src,
                "pass": {"kernels": " + ".join(kk["kernel"] for kk in kernels), "ms": round(k_ms, 3), "algorithmic_bytes": alg_bytes,
                         "achieved": round(alg_bytes / (k_ms * 1e-3) / 1e9, 1), "frac": round(alg_bytes / (k_ms * 1e-3) / 1e9 / peak, 4),
                         "traffic": traffic.get("pass")},
                "kernels": kernels}

    # e2e: host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        ne = int(min(n, args.e2e_reads))
        ue = max(1, int(ne / (n / u)))
        eb = sq.synth_reads(ne, ue, L, L, seed=SEED, first_read=rank * ne)
        h_ascii = torch.empty(ne * L, dtype=torch.uint8).pin_memory()
        h_ascii.copy_(eb.ascii[: ne * L])
        h_lens = torch.full((ne,), L, dtype=torch.uint8).pin_memory()      # one length per read, as a list of bytes carries
        h_words = torch.empty((ne,) if W == 1 else (ne, 3), dtype=torch.int64).pin_memory()
        del eb
        ectr = sq.DeviceCounter(klass, expected_unique=ue)
        eowner = sq.DeviceCounter(klass, expected_unique=int(1.1 * ue / world) + 1024, hash_rot=world.bit_length() - 1) if world > 1 else None
        rep = _lib.Report()

        def estep():
            _lib.check(lib.ssq_counter_clear(ectr.handle))
            _lib.check(lib.ssq_host_pack_count_lens(h, ectr.handle, h_ascii.data_ptr(), h_lens.data_ptr(), ne, h_words.data_ptr(),
                                                    1 << 22, C.byref(rep)))
            assert rep.code == 0
            if world > 1:
                _lib.check(lib.ssq_counter_clear(eowner.handle))
                if state["exchange"] is not None:
                    merge_peer(ectr, eowner, state["exchange"])
                else:
                    merge_alltoall(ectr, owner=eowner)
                return len(eowner)
            return len(ectr)

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
        e2e = {"value": round(world * ne * L / (e_ms * 1e-3) / 1e9, 3), "unit": "Gbases/s",
               "h2d_bytes_per_step": ne * L + ne, "d2h_bytes_per_step": ne * 8 * W + 8,
               "reads_per_step": ne, "ms_per_step": round(e_ms, 3),
               "call": "ssq_host_pack_count_lens (pinned host ASCII + one uint8 length per read in, packed words out, "
                       "4M-read chunks, H2D / kernel / D2H overlapped on three streams)"}

    if rank == 0:
        line = {
            "metric": "Gbases/s pack+count", "value": round(value, 2), "unit": "Gbases/s", "n_gpus": world, "steps": K,
            "warmup": WU, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_of(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "uniques": int(uniques_seen[0]),
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        if state["exchange"] is not None:
            state["exchange"].close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
