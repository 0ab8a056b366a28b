"""ShardedLocalCheckpointManager host logic on gloo (world sizes 2 and 4): striped fragments, coverage rule, rebuild of a
member that lost its storage."""
import pytest
import torch
import torch.distributed as dist

from _mp import run_ranks


def make_sd(seed):
    g = torch.Generator().manual_seed(seed)
    return {"w": torch.randn(301 + seed % 7, 9, generator=g), "opt": [{"m": torch.randn(33, generator=g).to(torch.bfloat16), "step": torch.tensor(float(seed))}],
            "ids": torch.randint(0, 10, (5 + seed % 3,), generator=g), "note": f"seed{seed}", "empty": torch.empty(0, 2)}


def same(a, b):
    ta, tb = list(a.tensors), list(b.tensors)
    return len(ta) == len(tb) and all(x.dtype == y.dtype and x.shape == y.shape and torch.equal(x.cpu(), y.cpu()) for x, y in zip(ta, tb)) and a.state_dict["note"] == b.state_dict["note"]


def _w(rank, world, root, jump, factor, kill, is_async):
    from _cpu_tasd import CpuTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.sharded_local_manager import (
        ShardedLocalCheckpointManager,
        fragment_range,
        shard_bytes_for,
    )

    mgr = ShardedLocalCheckpointManager.from_replication_params(root, replication_jump=jump, replication_factor=factor)
    members = mgr._members
    n = len(members) - 1
    q = AsyncCallsQueue(persistent=False)
    for it in (1, 2):
        req = mgr.save(CpuTensorAwareStateDict(make_sd(100 * it + rank)), it, is_async=is_async)
        if is_async:
            q.schedule_async_request(req)
            q.maybe_finalize_async_calls(blocking=True, no_dist=False)
        else:
            assert req is None
    import time

    time.sleep(0.5)  # async cleanup of iteration 1
    files = sorted(p.name for p in mgr.local_ckpt_dir.iterdir())
    want = [f"iter_0000002_{rank}_local.pt"] + [f"iter_0000002_{m}_local.s{mgr._others(m).index(rank)}of{n}.pt" for m in members if m != rank]
    assert files == sorted(want), (files, want)
    # fragments are byte ranges of the owner's packed buffer
    assert shard_bytes_for(1000, 3) == 512 and fragment_range(1000, 512, 1) == (512, 1000) and fragment_range(1000, 512, 2) == (1000, 1000)
    dist.barrier()
    if rank in kill:
        for p in mgr.local_ckpt_dir.iterdir():
            p.unlink()
    dist.barrier()
    mgr2 = ShardedLocalCheckpointManager(root, clique=mgr.clique)
    assert mgr2.find_latest() == 2
    loaded, cid = mgr2.load()
    assert cid == (2, rank, "") and same(loaded, CpuTensorAwareStateDict(make_sd(200 + rank)))
    dist.barrier()
    # two members of one clique lost: the iteration is no longer complete (striping tolerates ONE loss per clique)
    if len(members) >= 3:
        if rank in members[:2]:
            for p in mgr.local_ckpt_dir.iterdir():
                p.unlink()
        dist.barrier()
        assert ShardedLocalCheckpointManager(root, clique=mgr.clique).find_latest() == -1
    q.close()


@pytest.mark.parametrize("world,jump,factor,kill,is_async", [(2, 1, 2, (1,), True), (4, 1, 4, (2,), True), (4, 2, 2, (0, 3), False), (4, 1, 4, (0,), False)])
def test_sharded_save_and_rebuild(tmp_path, world, jump, factor, kill, is_async):
    run_ranks(_w, world, str(tmp_path), jump, factor, kill, is_async)
