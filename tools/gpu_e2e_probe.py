"""End-to-end probes on one GPU: sequential pack+drain vs the pipelined nvrx_snapshot, chunk sizes, 2-stream drain,
host-side overhead of the API call.  Prints a table; run under gpurun."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nvidia-resiliency-ext_b200"))
import torch  # noqa: E402

from bench import flatten, llama3_8b_shard_state  # noqa: E402
from nvidia_resiliency_ext.checkpointing.b200 import _cabi  # noqa: E402
from nvidia_resiliency_ext.checkpointing.b200.engine import DeviceBuffer, Event, HostBuffer, Plan, Stream  # noqa: E402

def main():
    lib = _cabi.lib()
    sd, total = llama3_8b_shard_state(torch.device("cuda"))
    tensors = flatten(sd)
    plan = Plan([t.data_ptr() for t in tensors], [t.numel() * 4 for t in tensors], None, device=0)
    stg = DeviceBuffer(plan.staging_bytes, 0)
    t0 = time.time()
    hb = HostBuffer.create(plan.staging_bytes, name=f"/nvrx_probe_{os.getpid()}", pin=True, device=0, prefault_threads=16)
    print(f"slot create+prefault(NUMA-local)+pin: {time.time()-t0:.2f} s", flush=True)
    cur = torch.cuda.current_stream().cuda_stream
    side, side2 = Stream(0), Stream(0)
    done, done2 = Event(0), Event(0)
    S = plan.staging_bytes


    def wall(fn, n=3):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t)
        ts.sort()
        return ts[len(ts) // 2]


    def sequential():
        plan.pack(stg.ptr, cur)
        ev = Event(0)
        ev.record(cur)
        side.wait_event(ev)
        _cabi.check(lib.nvrx_drain(hb.data_ptr, stg.ptr, S, 256 << 20, hb.progress_ptr, 0, side.handle, done.handle), "drain")
        done.synchronize()


    print(f"sequential pack + drain(256 MiB chunks): {wall(sequential)*1e3:8.2f} ms", flush=True)
    for mb in (64, 256, 1024):
        def pipelined(mb=mb):
            _cabi.check(lib.nvrx_snapshot(plan._h, stg.ptr, hb.data_ptr, mb << 20, hb.progress_ptr, 0, cur, side.handle, None, done.handle), "snap")
            done.synchronize()

        w = wall(pipelined)
        print(f"pipelined nvrx_snapshot chunk={mb:5d} MiB:   {w*1e3:8.2f} ms   ({total/w/1e9:.2f} GB/s)", flush=True)


    def two_stream():
        plan.pack(stg.ptr, cur)
        ev = Event(0)
        ev.record(cur)
        side.wait_event(ev)
        side2.wait_event(ev)
        half = (S // 2 + 511) // 512 * 512
        _cabi.check(lib.nvrx_drain(hb.data_ptr, stg.ptr, half, 0, None, 0, side.handle, done.handle), "drain")
        _cabi.check(lib.nvrx_drain(hb.data_ptr + half, stg.ptr + half, S - half, 0, None, 0, side2.handle, done2.handle), "drain")
        done.synchronize()
        done2.synchronize()


    print(f"pack + drain split over 2 streams:        {wall(two_stream)*1e3:8.2f} ms", flush=True)

    # host-side cost of the API call (time until async_save returns; GPU work enqueued, nothing waited for)
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=0, world_size=1)
    hb.close()
    stg.close()
    ck = TorchAsyncCheckpoint()
    for i in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        ck.async_save(sd, f"/dev/shm/nvrx_probe_{os.getpid()}.pt")
        t_call = time.perf_counter() - t
        snap = next(iter(ck._pending.values()))
        snap.wait()
        t_host = time.perf_counter() - t
        ck.finalize_async_save(blocking=True)
        t_persist = time.perf_counter() - t
        print(f"[{time.strftime('%X')}] async_save #{i}: call returns {t_call*1e3:7.2f} ms, host-safe {t_host*1e3:7.2f} ms ({total/t_host/1e9:.2f} GB/s), persisted {t_persist:6.2f} s ({total/t_persist/1e9:.2f} GB/s)", flush=True)
    ck.close()
    os.unlink(f"/dev/shm/nvrx_probe_{os.getpid()}.pt")
    dist.destroy_process_group()


if __name__ == "__main__":  # the persistent checkpoint worker is spawned: it re-imports this module
    main()
