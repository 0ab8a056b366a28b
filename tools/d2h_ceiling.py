"""What bounds the end-to-end snapshot rate at N GPUs?  N concurrent device -> host copies of 16 GB each, one process per GPU
(torchrun), with different kinds of host memory:

    cudahostalloc        torch pinned tensor (cudaHostAlloc), allocated by the rank's main thread wherever it runs
    cudahostalloc+bind   same, the process bound to the CPUs of the GPU's NUMA node before allocating
    slot                 the engine's snapshot slot: POSIX shm, first-touched on the GPU's NUMA node, cudaHostRegister
    slot-nonuma          the same without the NUMA-local first touch (NVRX_B200_NO_NUMA=1)

Every variant copies the same bytes with cudaMemcpyAsync in 256 MiB chunks on one stream; all ranks start together; the
per-rank rate is bytes / (max over ranks of the elapsed time).  Prints one JSON line per variant (rank 0).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_ceiling.py [--gb 16]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nvidia-resiliency-ext_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def numa_cpus(local):
    bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
        bdf = bdf.lower()[-12:]
    except Exception:  # noqa: BLE001
        return None, None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:  # noqa: BLE001
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=16.0)
    ap.add_argument("--variants", default="cudahostalloc,cudahostalloc+bind,slot,slot-nonuma")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi
    from nvidia_resiliency_ext.checkpointing.b200.engine import Event, HostBuffer, Stream

    n = int(args.gb * 1e9) // 4096 * 4096
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    src.random_(0, 255)
    node, cpus = numa_cpus(local)
    all_cpus = os.sched_getaffinity(0)
    lib = _cabi.lib()
    side = Stream(local)
    done = Event(local)
    for variant in args.variants.split(","):
        os.sched_setaffinity(0, all_cpus)
        os.environ.pop("NVRX_B200_NO_NUMA", None)
        hb = pinned = None
        t_alloc = time.perf_counter()
        if variant.startswith("cudahostalloc"):
            if variant.endswith("+bind") and cpus:
                os.sched_setaffinity(0, cpus)
            pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
            dst = pinned.data_ptr()
        else:
            if variant == "slot-nonuma":
                os.environ["NVRX_B200_NO_NUMA"] = "1"
            hb = HostBuffer.create(n, name=f"/nvrx_b200_ceil_{os.getpid()}_{variant.replace('-', '')}", pin=True, device=local, prefault_threads=16)
            dst = hb.data_ptr
        t_alloc = time.perf_counter() - t_alloc
        rates = []
        for rep in range(args.reps + 1):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _cabi.check(lib.nvrx_drain(dst, src.data_ptr(), n, 256 << 20, None, 0, side.handle, done.handle), "nvrx_drain")
            done.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rep:
                rates.append(n / dt.item() / 1e9)
        allocs = torch.tensor([t_alloc], dtype=torch.float64, device="cuda")
        dist.all_reduce(allocs, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"variant": variant, "n_gpus": world, "GB_per_rank": round(n / 1e9, 2), "per_gpu_GBps_median": round(sorted(rates)[len(rates) // 2], 2),
                              "per_gpu_GBps_all": [round(r, 2) for r in rates], "aggregate_GBps": round(world * sorted(rates)[len(rates) // 2], 1),
                              "alloc_s_max": round(allocs.item(), 2), "gpu0_numa_node": node, "host_cpus": os.cpu_count()}), flush=True)
        if hb is not None:
            hb.close()
        del pinned
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
