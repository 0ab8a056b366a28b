"""``bench.py --config c4``: BASELINE configs C4 / C5 -- replicated local checkpoints of the 16 GB/rank state through the
manager API, and the restore of a rank that lost its storage.  Both arms, every rank, same state, same box.

Layouts (``--layouts``, default all that fit the box):

  pairs     ``LocalCheckpointManager`` + ``CliqueReplicationStrategy.from_replication_params(world/2, 2)``: reference
            semantics (every clique member stores every member's shard), cliques of 2, FULL 16 GB/rank.  This is the layout
            both arms can run at the stated size (full replication with F=8 needs 128 GB of pinned host per rank
            in either implementation, more than the box has) -> the headline ``value``.
  full      same with ONE clique of all ranks (F = world), state shrunk by 1/world so that F x S fits (SURVEY 8d C4 note).
  striped   ``ShardedLocalCheckpointManager`` (F = world): the north star's all-to-all layout, FULL 16 GB/rank -- the packed
            snapshot is cut into F-1 fragments, one per peer.  The reference has no such layout (engine arm only).

Per layout: ``stall_ms`` (wall time ``save()`` keeps the training stream blocked), ``host_safe_s`` (replicas in pinned host
memory), ``persist_s`` (files complete, ``maybe_finalize_async_calls(blocking=True)``), ``restore_s`` (rank 1 deletes its
directory; ``find_latest() + load()`` on every rank), ``bit_exact`` (every restored tensor of every rank against the live
state, all bytes), NVLink bytes a rank sends and the rate during the stall.

Reference arm = ``oracle/reference_port.py``: ``reference_replicated_save`` (F x N ``dist.broadcast`` + per-tensor D2H + sync +
fork + ``torch.save``) and the per-tensor retrieve (``torch.load`` + N x H2D on the holder, N x ``dist.send`` / ``dist.recv``).
"""
import json
import os
import shutil
import time
from pathlib import Path

import torch
import torch.distributed as dist

import bench as B


def _exact(got, want):
    if len(got) != len(want):
        return False
    for a, b in zip(got, want):
        if not (a.is_cuda and B.bits_equal(a, b)):
            return False
    return True


def _all_ok(ok: bool) -> bool:
    t = torch.tensor([int(bool(ok))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def _root(kind):
    return Path("/dev/shm") / f"nvrx_b200_c4_{os.environ.get('MASTER_PORT', '0')}_{kind}"


def engine_layout(kind, rank, world, local, sd, total, tensors, iters, warm):
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.sharded_local_manager import ShardedLocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    root = _root(kind)
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    dist.barrier()
    strat = None
    if kind == "striped":
        factor = world
        mgr = ShardedLocalCheckpointManager.from_replication_params(root, replication_jump=1, replication_factor=world)
        nv_bytes = total
    else:
        factor = 2 if kind == "pairs" else world
        jump = world // 2 if kind == "pairs" else 1
        strat = CliqueReplicationStrategy.from_replication_params(jump, factor)
        mgr = LocalCheckpointManager(root, repl_strategy=strat)
        nv_bytes = (factor - 1) * total
    engine = SnapshotEngine.get(local)
    q = AsyncCallsQueue(persistent=False)
    ev = torch.cuda.Event()
    rows = []
    for it in range(1, warm + iters + 1):
        tasd = BasicTensorAwareStateDict(B.fresh_containers(sd))
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        req = mgr.save(tasd, it, is_async=True)
        ev.record()
        ev.synchronize()
        t1 = time.perf_counter()
        for s in (engine._side, getattr(engine, "_comm", None)):
            if s is not None:
                s.synchronize()  # drains (and the streamed exchange) are done: every replica is in pinned host memory
        t2 = time.perf_counter()
        q.schedule_async_request(req)
        q.maybe_finalize_async_calls(blocking=True, no_dist=False)
        t3 = time.perf_counter()
        if it > warm:
            rows.append((t1 - t0, t2 - t0, t3 - t0))
        del tasd, req
    mode = getattr(engine, "last_exchange", "?")
    stall, safe, persist = (B.max_over_ranks(B.mean([r[i] for r in rows])) for i in range(3))
    # C5: one rank loses its storage, gets its shard back from the replica holder(s) and scatters it
    dist.barrier()
    time.sleep(1.0)  # the background cleanup of the previous iteration
    lost = 1 % world
    if rank == lost:
        for p in list(mgr.local_ckpt_dir.iterdir()):
            p.unlink(missing_ok=True)
    dist.barrier()
    mgr2 = ShardedLocalCheckpointManager(root, clique=mgr.clique) if kind == "striped" else LocalCheckpointManager(root, repl_strategy=strat)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    latest = mgr2.find_latest()
    loaded, cid = mgr2.load()
    torch.cuda.synchronize()
    restore = B.max_over_ranks(time.perf_counter() - t0)
    ok = latest == warm + iters and tuple(cid[:2]) == (warm + iters, rank) and _exact(list(loaded.tensors), tensors)
    ok = _all_ok(ok)
    del loaded
    q.close()
    engine.trim()
    dist.barrier()
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    dist.barrier()
    return {
        "manager": "ShardedLocalCheckpointManager" if kind == "striped" else "LocalCheckpointManager + CliqueReplicationStrategy",
        "factor": factor, "state_bytes_per_rank": total, "exchange": mode,
        "stall_ms": round(stall * 1e3, 2), "host_safe_s": round(safe, 3), "persist_s": round(persist, 3), "restore_s": round(restore, 3),
        "bit_exact": ok, "lost_rank": lost,
        "nvlink_out_bytes_per_gpu": nv_bytes, "nvlink_GBps_per_gpu_during_stall": round(nv_bytes / stall / 1e9, 1),
        "snapshot_GBps_all_ranks": round(world * total / stall / 1e9, 1),
        "nccl_payload_ops": 0 if mode == "p2p-fused" else 1,
    }


def reference_layout(kind, rank, world, local, sd, total, tensors, iters, warm):
    from oracle import reference_port as rp

    root = _root("ref_" + kind)
    mydir = root / str(rank)
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    dist.barrier()
    mydir.mkdir(parents=True, exist_ok=True)
    factor = 2 if kind == "pairs" else world
    jump = world // 2 if kind == "pairs" else 1
    # cliques n, n+J, ..., n+(F-1)J  (local/replication/group_utils.py:120-146)
    cliques = []
    for base in range(0, world, jump * factor):
        for off in range(jump):
            cliques.append([base + off + k * jump for k in range(factor)])
    group, members = None, None
    for c in cliques:
        g = dist.new_group(c)
        if rank in c:
            group, members = g, c
    rows = []
    for it in range(1, warm + iters + 1):
        paths = [mydir / f"iter_{it:07d}_{m}_local.pt" for m in members]
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        res = rp.reference_replicated_save(B.fresh_containers(sd), paths, group)
        if it > warm:
            rows.append((res["stall"], res["stall"], res["total"]))
        for p in mydir.glob(f"iter_{it - 1:07d}_*"):
            p.unlink()
    stall, safe, persist = (B.max_over_ranks(B.mean([r[i] for r in rows])) for i in range(3))
    last = warm + iters
    lost = 1 % world
    if rank == lost:
        for p in list(mydir.iterdir()):
            p.unlink()
    dist.barrier()
    # retrieve (base_manager.py:206-234, strategies.py:143-179, group_utils.py:378-449): holders _load_fn what they send
    # (torch.load + N x H2D), then N x send / recv; everybody else loads its own file
    lost_clique = next(c for c in cliques if lost in c)
    holder = sorted(m for m in lost_clique if m != lost)[0]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    if rank == lost:
        got = [torch.empty_like(t) for t in tensors]  # hollow.init_tensors(): placeholders remember shape/dtype/device
        for t in got:
            dist.recv(t, holder)
    else:
        mine = torch.load(mydir / f"iter_{last:07d}_{rank}_local.pt", weights_only=False)
        got = [t.to("cuda", non_blocking=False) for t in mine]
        if rank == holder:
            theirs = torch.load(mydir / f"iter_{last:07d}_{lost}_local.pt", weights_only=False)
            for t in theirs:
                dist.send(t.to("cuda", non_blocking=False).cuda(), lost)
    torch.cuda.synchronize()
    restore = B.max_over_ranks(time.perf_counter() - t0)
    ok = _all_ok(_exact(got, tensors))  # every rank ends with its OWN state on its GPU (the lost one from the replica)
    del got
    dist.barrier()
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    dist.barrier()
    return {
        "manager": "reference flow: CliqueReplicationStrategy.replicate (F x N broadcasts + per-tensor D2H) + fork + torch.save",
        "factor": factor, "state_bytes_per_rank": total, "exchange": f"{factor} x {len(tensors)} dist.broadcast",
        "stall_ms": round(stall * 1e3, 2), "host_safe_s": round(safe, 3), "persist_s": round(persist, 3), "restore_s": round(restore, 3),
        "bit_exact": ok, "lost_rank": lost,
        "nvlink_out_bytes_per_gpu": (factor - 1) * total, "nvlink_GBps_per_gpu_during_stall": round((factor - 1) * total / stall / 1e9, 1),
        "snapshot_GBps_all_ranks": round(world * total / stall / 1e9, 1),
        "nccl_payload_ops": factor * len(tensors),
    }


def run(args, rest, rank, world, local):
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="pairs,striped,full")
    ap.add_argument("--c4-iters", type=int, default=2)
    ap.add_argument("--c4-warm", type=int, default=1)
    c4, _ = ap.parse_known_args(rest)
    if world < 2:
        raise SystemExit("--config c4 needs at least 2 GPUs (torchrun --nproc-per-node N bench.py --config c4 --gpus N)")
    dev = torch.device("cuda", local)
    engine_arm = args.impl == "engine"
    fn = engine_layout if engine_arm else reference_layout
    clocks = B.ClockSampler(local)
    clocks.__enter__()
    out = {}
    for kind in [k for k in c4.layouts.split(",") if k]:
        if kind == "striped" and not engine_arm:
            continue  # the reference has no striped layout
        if kind == "full" and world == 2:
            continue  # identical to pairs
        scale = args.scale * (1.0 / world if kind == "full" else 1.0)
        sd, total = B.llama3_8b_shard_state(dev, seed=1234 + rank, scale=scale)
        tensors = B.flatten(sd)
        out[kind] = fn(kind, rank, world, local, sd, total, tensors, c4.c4_iters, c4.c4_warm)
        out[kind]["scale"] = scale
        del sd, tensors
        torch.cuda.empty_cache()
    clocks.__exit__(None, None, None)
    head = out.get("pairs") or next(iter(out.values()))
    line = {
        "metric": B.METRIC, "value": head["snapshot_GBps_all_ranks"], "unit": "GB/s", "n_gpus": world,
        "steps": c4.c4_iters, "warmup": c4.c4_warm, "ms_per_step": head["stall_ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": args.impl,
        "value_definition": "state bytes of all ranks / training-stream stall of LocalCheckpointManager.save(is_async=True) with "
                            "clique replication (layout `pairs`, full size); other layouts under `layouts`",
        "config": {"workload": f"C4/C5: LocalCheckpointManager with clique replication on {world} GPUs, 16.06 GB Llama-3-8B-shaped state "
                               "per rank, replicas to pinned host + files on /dev/shm, then restore of a rank that lost its storage",
                   "scale": args.scale, "layouts": list(out)},
        "e2e": {"value": round(world * head["state_bytes_per_rank"] / head["host_safe_s"] / 1e9, 2), "unit": "GB/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(head["factor"] * head["state_bytes_per_rank"]),
                "definition": "state bytes of all ranks / wall time from save() until every replica is in pinned host memory"},
        "stall_ms": head["stall_ms"], "persist_s": head["persist_s"], "restore_s": head["restore_s"],
        "layouts": out, "gpu_launches": (c4.c4_iters + c4.c4_warm + 1) * len(out) if engine_arm else 0,
        "clocks": clocks.summary(), "host_cores_on_box": os.cpu_count(),
    }
    if not engine_arm:
        line["cpu_baseline"] = {"value": line["value"], "unit": "GB/s", "cores": 1, "kind": "port",
                                "sample": "every step, every rank: the reference's replicate + D2H + sync (full state); torch.save child on 1 core per rank"}
        line["e2e"]["d2h_bytes_per_step"] = 0
    return line


if __name__ == "__main__":
    raise SystemExit("run as: bench.py --config c4")
