"""C4 / C5 measurement: LocalCheckpointManager.save with clique replication + restore of a lost shard (torchrun, 1 rank/GPU).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_replicate.py [--scale 0.25] [--mode p2p|nccl]

Prints one JSON line (rank 0): GPU-side stall of save() (barrier + fused pack/all-gather + barrier, CUDA events, max
over ranks), NVLink bytes and GB/s per GPU, time until all replicas are in host memory, restore time of a lost shard."""
import argparse
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

from bench import flatten, llama3_8b_shard_state, max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--factor", type=int, default=0, help="replication factor (default: world size)")
    ap.add_argument("--kernel-only", action="store_true",
                    help="time only the fused pack+exchange kernels at the given scale (no host slots, no files)")
    ap.add_argument("--layout", default="full", choices=["full", "sharded"],
                    help="full = reference semantics (every member stores every member's shard); sharded = striped fragments (all-to-all)")
    args = ap.parse_args()
    os.environ["NVRX_B200_EXCHANGE"] = args.mode
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    factor = args.factor or world
    if args.kernel_only:
        return kernel_only(args, rank, world, local)
    sd, total = llama3_8b_shard_state(torch.device("cuda", local), seed=1234 + rank, scale=args.scale)
    want = [t.clone() for t in flatten(sd)] if args.scale <= 0.3 else None
    root = Path("/dev/shm") / f"nvrx_b200_repl_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    dist.barrier()
    if args.layout == "sharded":
        from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.sharded_local_manager import ShardedLocalCheckpointManager

        mgr = ShardedLocalCheckpointManager.from_replication_params(root, replication_jump=1, replication_factor=factor)
        strat = None
    else:
        strat = CliqueReplicationStrategy.from_replication_params(1, factor)
        mgr = LocalCheckpointManager(root, repl_strategy=strat)
    q = AsyncCallsQueue(persistent=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = []
    for it in range(1, args.iters + 1):
        tasd = BasicTensorAwareStateDict(llama_like(sd))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        e0.record()
        req = mgr.save(tasd, it, is_async=True)
        e1.record()
        t_call = time.perf_counter() - t0
        e1.synchronize()
        t_stall = time.perf_counter() - t0
        gpu_ms = e0.elapsed_time(e1)
        q.schedule_async_request(req)
        # replicas in host memory == drain of the exchange buffer done (the request's finalize releases the handles)
        q.maybe_finalize_async_calls(blocking=True, no_dist=False)
        t_done = time.perf_counter() - t0
        rows.append((t_call, t_stall, gpu_ms, t_done))
    mode = SnapshotEngine.get().last_exchange
    t_call, t_stall, gpu_ms, t_done = (max_over_ranks(sorted(r[i] for r in rows)[len(rows) // 2]) for i in range(4))
    # C5: rank 1 loses its storage, gets its shard back from a replica holder and scatters it
    dist.barrier()
    time.sleep(1.0)  # let the background cleanup of the previous iteration finish
    if rank == 1 % world:
        for p in list(mgr.local_ckpt_dir.iterdir()):
            p.unlink(missing_ok=True)
    dist.barrier()
    if args.layout == "sharded":
        mgr2 = ShardedLocalCheckpointManager(root, clique=mgr.clique)
    else:
        mgr2 = LocalCheckpointManager(root, repl_strategy=strat)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    latest = mgr2.find_latest()
    loaded, cid = mgr2.load()
    torch.cuda.synchronize()
    t_restore = max_over_ranks(time.perf_counter() - t0)
    ok = latest == args.iters and cid == (args.iters, rank, "")
    if want is not None:
        got = list(loaded.tensors)
        ok = ok and len(got) == len(want) and all(torch.equal(a.view(-1).view(torch.uint8), b.view(-1).view(torch.uint8)) for a, b in zip(got, want))
    okt = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    q.close()
    dist.barrier()
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        nv_bytes = total if args.layout == "sharded" else (factor - 1) * total
        print(json.dumps({
            "config": f"C4/C5 [{args.layout}]: {'ShardedLocalCheckpointManager' if args.layout == 'sharded' else 'LocalCheckpointManager + CliqueReplicationStrategy'}"
                      f"(J=1, F={factor}) on {world} GPUs, {total/1e9:.2f} GB/rank (scale {args.scale})",
            "exchange": mode, "save_call_ms": round(t_call * 1e3, 2), "stall_ms_wall": round(t_stall * 1e3, 2), "stall_ms_gpu": round(gpu_ms, 2),
            "nvlink_out_bytes_per_gpu": nv_bytes, "nvlink_GBps_per_gpu_during_stall": round(nv_bytes / (gpu_ms * 1e-3) / 1e9, 1),
            "replicas_persisted_s": round(t_done, 2), "restore_lost_shard_s": round(t_restore, 2), "bit_exact": bool(okt.item()),
            "nccl_ops_on_data_path": 0 if mode == "p2p-fused" else 1, "reference_nccl_ops": f"{factor} x 1455 broadcasts",
        }), flush=True)
    dist.destroy_process_group()


def kernel_only(args, rank, world, local):
    """C4 at full size without touching host memory: the fused kernels over NVLink, CUDA-event timed, max over ranks."""
    from nvidia_resiliency_ext.checkpointing.b200 import exchange as xch
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.sharded_local_manager import shard_bytes_for
    from nvidia_resiliency_ext.checkpointing.local.replication.group_utils import GroupWrapper

    sd, total = llama3_8b_shard_state(torch.device("cuda", local), seed=1234 + rank, scale=args.scale)
    tensors = flatten(sd)
    engine = SnapshotEngine.get(local)
    plan = engine._plan_for(tensors, [False] * len(tensors))
    S = plan.staging_bytes
    grp = GroupWrapper()
    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {"config": f"C4 kernel-only on {world} GPUs, {total/1e9:.2f} GB/rank (scale {args.scale})", "mode": args.mode}

    def timed(fn):
        ms = []
        for it in range(2 + args.iters):
            xch._clique_barrier(engine, grp)
            e0.record()
            fn()
            e1.record()
            xch._clique_barrier(engine, grp)
            e1.synchronize()
            if it >= 2:
                ms.append(e0.elapsed_time(e1))
        return max_over_ranks(sorted(ms)[len(ms) // 2])

    staging = engine._ensure_staging(S)
    if world > 1:
        # sharded: own full copy + fragment k -> k-th other member
        n = world - 1
        sb = shard_bytes_for(S, n)
        xbuf, bases = xch.shared_exchange(engine, grp, n * sb)
        if bases is not None:
            others = [r for r in range(world) if r != rank]
            dest = [bases[m] + [r for r in range(world) if r != m].index(rank) * sb for m in others]
            t = timed(lambda: plan.pack_sharded(staging.ptr, dest, sb, 0, stream))
            out["sharded_fused_ms"] = round(t, 3)
            out["sharded_nvlink_GBps_per_gpu"] = round(S / (t * 1e-3) / 1e9, 1)
            out["sharded_hbm_GBps"] = round(3 * S / (t * 1e-3) / 1e9, 1)  # read S, write S local, S remote
        whole = xch.as_uint8_tensor(staging.ptr, S, local)
        recv = xch.as_uint8_tensor(xbuf.ptr, n * sb, local)
        def frag(k):
            return min(S, (k + 1) * sb) - min(S, k * sb)

        def idx(owner, holder):  # which fragment of `owner` lives on `holder`
            return [r for r in range(world) if r != owner].index(holder)

        splits_in = [0 if q == rank else frag(idx(rank, q)) for q in range(world)]
        splits_out = [0 if q == rank else frag(idx(q, rank)) for q in range(world)]

        def nccl_a2a():
            plan.pack(staging.ptr, stream)
            torch.distributed.all_to_all_single(recv[: sum(splits_out)], whole[: sum(splits_in)], splits_out, splits_in)
        t = timed(nccl_a2a)
        out["sharded_pack_plus_nccl_alltoall_ms"] = round(t, 3)
        # full replication: every member gets every member's packed snapshot
        free, _ = torch.cuda.mem_get_info()
        if free > world * S + (8 << 30):
            xbuf, bases = xch.shared_exchange(engine, grp, world * S)
            if bases is not None:
                t = timed(lambda: plan.pack_broadcast(bases, rank * S, stream))
                out["full_fused_ms"] = round(t, 3)
                out["full_nvlink_GBps_per_gpu"] = round((world - 1) * S / (t * 1e-3) / 1e9, 1)
            big = xch.as_uint8_tensor(xbuf.ptr, world * S, local)
            def nccl_ag():
                plan.pack(xbuf.ptr + rank * S, stream)
                torch.distributed.all_gather_into_tensor(big, big[rank * S : (rank + 1) * S])
            t = timed(nccl_ag)
            out["full_pack_plus_nccl_allgather_ms"] = round(t, 3)
    t = timed(lambda: plan.pack(staging.ptr, stream))
    out["local_pack_ms"] = round(t, 3)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def llama_like(sd):
    """Fresh nested containers around the same tensors (save() hollows / rewrites the container it is given)."""
    return {"model": dict(sd["model"]), "optimizer": {"state": {k: dict(v) for k, v in sd["optimizer"]["state"].items()}}}


if __name__ == "__main__":
    main()
