"""Where does a restore of the 16 GB state go?  Saves the C2 state once through LocalCheckpointManager, then times the stages
of the default restore separately (torch.load(mmap) / offsets / file->ring->device pipeline for several thread counts and
chunk sizes / scatter), the page-cache read rate alone (pread into a pinned buffer without H2D) and the plain H2D rate."""
import json
import os
import shutil
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nvidia-resiliency-ext_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import flatten, fresh_containers, llama3_8b_shard_state  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    torch.cuda.set_device(0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi, ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine, _u64_array
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    sd, total = llama3_8b_shard_state(torch.device("cuda", 0), scale=scale)
    root = Path("/dev/shm") / f"nvrx_b200_rb_{os.getpid()}"
    mgr = LocalCheckpointManager(root)
    mgr.save(BasicTensorAwareStateDict(fresh_containers(sd)), 1, is_async=False)
    path = mgr._local_ckpt_path_from_id(mgr._ckpt_id(1))
    engine = SnapshotEngine.get(0)
    out = {"state_GB": round(total / 1e9, 2)}
    try:
        t0 = time.perf_counter()
        loaded = torch.load(path, weights_only=False, mmap=True)
        out["torch_load_mmap_s"] = round(time.perf_counter() - t0, 3)
        host = list(loaded.tensors)
        t0 = time.perf_counter()
        offs = ptzip.tensor_offsets_in_file(path, host)
        out["offsets_s"] = round(time.perf_counter() - t0, 3)
        dev = [torch.empty(t.shape, dtype=t.dtype, device="cuda") for t in host]
        plan = engine._plan_for(dev, [False] * len(dev))
        staging = engine._ensure_staging(plan.staging_bytes)
        live = [(o, nb, fo) for o, nb, fo in zip(plan.offsets, plan.packed_nbytes, offs) if nb]
        stream = torch.cuda.current_stream().cuda_stream
        lib = _cabi.lib()
        fd = os.open(path, os.O_RDONLY)
        sweep = {}
        for threads, chunk_mb, ring in [(16, 64, 4), (32, 64, 4), (64, 64, 4), (32, 16, 8), (32, 256, 4), (8, 64, 4), (32, 64, 4)]:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _cabi.check(lib.nvrx_fill_from_fd(staging.ptr, plan.staging_bytes, fd, len(live), _u64_array([x[0] for x in live]),
                                              _u64_array([x[1] for x in live]), _u64_array([x[2] for x in live]), chunk_mb << 20, ring, threads, 0, stream),
                        "nvrx_fill_from_fd")
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sweep[f"t{threads}_c{chunk_mb}_r{ring}"] = {"s": round(dt, 3), "GBps": round(total / dt / 1e9, 1)}
        os.close(fd)
        out["fill_from_fd"] = sweep
        t0 = time.perf_counter()
        plan.scatter(staging.ptr, stream)
        torch.cuda.synchronize()
        out["scatter_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        ok = all(torch.equal(a, b) for a, b in list(zip(dev, flatten(sd)))[::40])
        out["bit_exact_sample"] = ok
        # the two ceilings of the pipeline
        pinned = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            d.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()
        out["h2d_pinned_GBps"] = round(8 * (1 << 30) / (time.perf_counter() - t0) / 1e9, 1)
        from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer

        hb = HostBuffer.create(4 << 30, name=f"/nvrx_b200_rbslot_{os.getpid()}", pin=True, device=0, prefault_threads=16)
        fd = os.open(path, os.O_RDONLY)
        for threads in (8, 16, 32, 64):
            t0 = time.perf_counter()
            hb.readv_fd([0], [4 << 30], [1 << 20], fd, threads=threads)
            out[f"pread_{threads}thr_GBps"] = round((4 << 30) / (time.perf_counter() - t0) / 1e9, 1)
        os.close(fd)
        hb.close()
        # the whole public call
        del loaded, host
        del dev
        torch.cuda.empty_cache()
        for label in ("first", "second"):
            mgr2 = LocalCheckpointManager(root)
            engine.trace = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mgr2.find_latest()
            t1 = time.perf_counter()
            back, _ = mgr2.load()
            torch.cuda.synchronize()
            out[f"find_latest_plus_load_s_{label}"] = round(time.perf_counter() - t0, 3)
            out[f"find_latest_s_{label}"] = round(t1 - t0, 3)
            out[f"load_stages_{label}"] = {k: round(v, 3) for k, v in engine.trace.get("restore", {}).items()}
            del back
    finally:
        shutil.rmtree(root, ignore_errors=True)
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
