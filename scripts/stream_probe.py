"""Development: the streamed exchange (ssq_comm_attach) at world = 1 on one GPU, stage by stage, with timings.
usage: stream_probe.py [reads] [distinct] [expected_unique]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler
import torch
import torch.distributed as dist


def say(*a):
    print(*a, flush=True)


def main():
    faulthandler.dump_traceback_later(int(os.environ.get("SSQ_TEST_HANG_S", "60")), exit=True)
    import shortseq_b200 as sq
    from shortseq_b200.distributed import Comm
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_500_000
    u = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_200_000
    eu = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2_000_000
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    b = sq.synth_reads(n, u, 32, 32, seed=0x5EED0091)
    local = sq.DeviceCounter(0, expected_unique=eu)
    owner = sq.DeviceCounter(0, expected_unique=eu)
    comm = Comm(local.ctx)
    say("attach ->", comm.attach(local, owner))
    for rep in range(3):
        local.clear(); owner.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        local.pack_count(b)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        comm.merge(local, owner)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        say(f"rep {rep}: pass {1e3 * (t1 - t0):.2f} ms, merge {1e3 * (t2 - t1):.2f} ms (exchange {comm.exchange_ms:.3f} merge {comm.merge_ms:.3f}) streamed={comm.last_streamed} "
            f"local={len(local)} owner={len(owner)}")
    say("small insert")
    local.insert(sq.pack_batch(sq.synth_reads(1000, u, 32, 32, seed=0x5EED0091), klass=0))
    say("inserted"); owner.clear()
    comm.merge(local, owner)
    say("unstreamed merge done: streamed =", comm.last_streamed, "owner", len(owner))
    comm.attach(None, None)
    say("detached")
    local.clear(); owner.clear(); local.pack_count(b); comm.merge(local, owner)
    say("after detach: streamed =", comm.last_streamed, "owner", len(owner))
    comm.close()
    dist.destroy_process_group()
    say("done")


if __name__ == "__main__":
    main()
