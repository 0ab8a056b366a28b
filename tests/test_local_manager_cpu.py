"""LocalCheckpointManager + clique replication host logic on CPU tensors (gloo; world sizes 1, 2 and 4).
Mirrors reference tests/checkpointing/unit/test_basic_local.py and pins ``replicate`` to the reference's own
output (tests/golden/replicate_2rank.json)."""
import hashlib
import json
import os
import time
from pathlib import Path

import pytest
import torch
import torch.distributed as dist

from _mp import run_ranks
from conftest import GOLDEN


def sha(t: torch.Tensor) -> str:
    c = t.detach().cpu().contiguous()
    raw = c.view(-1).view(torch.uint8).numpy().tobytes() if c.numel() else b""
    return hashlib.sha256(str(c.dtype).encode() + str(tuple(c.shape)).encode() + raw).hexdigest()


def cpu_tasd(sd):
    from _cpu_tasd import CpuTensorAwareStateDict

    return CpuTensorAwareStateDict(sd)


def make_sd(seed):
    g = torch.Generator().manual_seed(seed)
    return {"w": torch.randn(33, 9, generator=g), "opt": [{"m": torch.randn(33, 9, generator=g), "step": torch.tensor(float(seed))}],
            "ids": torch.randint(0, 10, (5,), generator=g), "note": f"seed{seed}"}


def same(a, b):
    ta, tb = list(a.tensors), list(b.tensors)
    return len(ta) == len(tb) and all(x.dtype == y.dtype and torch.equal(x.cpu(), y.cpu()) for x, y in zip(ta, tb)) and a.state_dict["note"] == b.state_dict["note"]


# ---- world of one ---------------------------------------------------------------------------------
@pytest.mark.parametrize("is_async", [False, True])
def test_basic_save_load_scenarios(tmp_path, dist_1rank, is_async):
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.base_manager import CheckpointingException
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    root = tmp_path / "subdir"  # does not exist yet

    def do_save(mgr, sd, it):
        req = mgr.save(sd, it, is_async)
        if is_async:
            req.execute_sync()
        else:
            assert req is None

    mgr = LocalCheckpointManager(root)
    with pytest.raises(CheckpointingException):
        mgr.load()  # find_latest not called
    with pytest.raises(CheckpointingException):
        mgr._ckpt_id(-1)
    sd1 = cpu_tasd(make_sd(1))
    do_save(mgr, sd1, 1)
    assert mgr.find_latest() == 1
    loaded, cid = mgr.load()
    assert same(loaded, cpu_tasd(make_sd(1))) and cid == (1, 0, "")

    mgr = LocalCheckpointManager(root)  # "restart"
    assert mgr.find_latest() == 1
    loaded, cid = mgr.load()
    assert same(loaded, cpu_tasd(make_sd(1)))

    mgr = LocalCheckpointManager(root)
    first = mgr._local_ckpt_path_from_id(mgr._ckpt_id(1))
    os.remove(first)
    assert mgr.find_latest() == -1

    do_save(mgr, cpu_tasd(make_sd(1)), 1)
    assert first.exists()
    do_save(mgr, cpu_tasd(make_sd(2)), 2)
    time.sleep(0.4)
    assert not first.exists()  # older iteration cleaned up
    assert mgr._local_ckpt_path_from_id(mgr._ckpt_id(2)).exists()
    with pytest.raises(AssertionError):
        mgr2 = LocalCheckpointManager(root)
        mgr2.find_latest()
        mgr2.save(cpu_tasd(make_sd(1)), 1)  # older than latest


def test_dirty_file_means_same_machine_replication(tmp_path, dist_1rank):
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.base_manager import SameMachineReplicationException
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(tmp_path)
    mgr._ensure_dir()
    mgr._local_ckpt_path_from_id(mgr._ckpt_id(3), True).touch()
    with pytest.raises(SameMachineReplicationException):
        mgr._save(cpu_tasd(make_sd(3)), mgr._ckpt_id(3))
    assert mgr._my_ckpt_ids() == []  # dirty files are invisible
    mgr._cleanup_failed_save(3)
    assert list(mgr.local_ckpt_dir.iterdir()) == []


def test_async_save_through_temporal_queue(tmp_path, dist_1rank):
    """The documented way to use the manager: AsyncCallsQueue(persistent=False) (examples/checkpointing/local_ckpt.py)."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(tmp_path)
    q = AsyncCallsQueue(persistent=False)
    req = mgr.save(cpu_tasd(make_sd(5)), 5, is_async=True)
    q.schedule_async_request(req)
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=False) == [0]
    assert mgr.find_latest() == 5
    loaded, _ = mgr.load()
    assert same(loaded, cpu_tasd(make_sd(5)))
    q.close()


# ---- world of two / four (spawned ranks, gloo) -------------------------------------------------------
def _w_find_latest_repl_disabled(rank, world, root, suffix):
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(root)
    mgr._ensure_dir()
    files = [mgr.local_ckpt_dir / mgr._filename_from_template(10, i, suffix) for i in range(world)]
    # rank 0: nothing; rank 1: ckpt_0 and ckpt_1; rank i: ckpt_i
    if rank == 1:
        files[0].touch()
    if rank != 0:
        files[rank].touch()
    assert mgr.find_latest() == -1
    dist.barrier()
    if rank == 0 and suffix == "":
        files[0].touch()
    dist.barrier()
    mgr.latest_iteration = -1
    assert mgr.find_latest() == (10 if suffix == "" else -1)


@pytest.mark.parametrize("suffix", ["", "some_suffix"])
def test_find_latest_replication_disabled(tmp_path, suffix):
    run_ranks(_w_find_latest_repl_disabled, 2, str(tmp_path), suffix)


def _w_replicate_golden(rank, world, out_dir):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", str(GOLDEN / "make_golden.py"))
    # only the seeded input generator of the golden script is needed (it guards its reference import by an assert)
    src = (GOLDEN / "make_golden.py").read_text()
    ns = {}
    start = src.index("def rank_tensors")
    end = src.index("def _replicate_worker")
    exec("import torch\n" + src[start:end], ns)  # noqa: S102 - test fixture code from this repo
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    tasd = cpu_tasd(ns["rank_tensors"](rank))
    strat = CliqueReplicationStrategy(dist.group.WORLD, target_device="cpu")
    got, ids = strat.replicate(tasd, (11, rank, ""))
    res = {"ids": [list(i) for i in ids], "tensors": [[sha(t) for t in sd.tensors] for sd in got],
           "tags": [sd.state_dict["tag"] for sd in got], "input_hollow": tasd.is_hollow}
    with open(Path(out_dir) / f"r{rank}.json", "w") as f:
        json.dump(res, f)


def test_replicate_matches_reference_output(tmp_path):
    run_ranks(_w_replicate_golden, 2, str(tmp_path))
    gold = json.load(open(GOLDEN / "replicate_2rank.json"))
    for r in range(2):
        mine = json.load(open(tmp_path / f"r{r}.json"))
        assert mine == gold[str(r)], f"rank {r}"


def _w_replicated_save_load(rank, world, root, jump, factor, kill):
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    strat = CliqueReplicationStrategy.from_replication_params(jump, factor)
    mgr = LocalCheckpointManager(root, repl_strategy=strat)
    q = AsyncCallsQueue(persistent=False)
    req = mgr.save(cpu_tasd(make_sd(100 + rank)), 3, is_async=True)
    q.schedule_async_request(req)
    q.maybe_finalize_async_calls(blocking=True, no_dist=False)
    files = sorted(p.name for p in mgr.local_ckpt_dir.iterdir())
    members = strat.local_group.ranks
    assert files == sorted(f"iter_0000003_{m}_local.pt" for m in members), files
    dist.barrier()
    if rank in kill:  # this rank loses its local storage
        for p in mgr.local_ckpt_dir.iterdir():
            p.unlink()
    dist.barrier()
    mgr2 = LocalCheckpointManager(root, repl_strategy=strat)
    assert mgr2.find_latest() == 3
    loaded, cid = mgr2.load()
    assert cid == (3, rank, "") and same(loaded, cpu_tasd(make_sd(100 + rank)))
    q.close()


@pytest.mark.parametrize("world,jump,factor,kill", [(2, 1, 2, (1,)), (4, 2, 2, (0, 3)), (4, 1, 4, (1, 2, 3))])
def test_replicated_save_and_retrieval(tmp_path, world, jump, factor, kill):
    run_ranks(_w_replicated_save_load, world, str(tmp_path), jump, factor, kill)


def _w_no_replica_left(rank, world, root):
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    strat = CliqueReplicationStrategy.from_replication_params(1, 2)
    mgr = LocalCheckpointManager(root, repl_strategy=strat)
    mgr.save(cpu_tasd(make_sd(rank)), 1, is_async=False)
    dist.barrier()
    for p in mgr.local_ckpt_dir.iterdir():  # every holder loses rank 0's shard
        if "_0_local" in p.name:
            p.unlink()
    dist.barrier()
    mgr.latest_iteration = -1
    assert mgr.find_latest() == -1  # iteration 1 is no longer covered for rank 0


def test_lost_everywhere_is_not_latest(tmp_path):
    run_ranks(_w_no_replica_left, 2, str(tmp_path))


def test_saving_does_not_switch_off_the_trainers_zip_checksums(tmp_path, dist_1rank, built_library, monkeypatch):
    """Checksums are on by default (nothing is switched).  With NVRX_B200_ZIP_CRC=0 the writer skips zip CRCs while it saves a
    snapshot; that is a process-wide PyTorch switch and a synchronous save runs in the trainer, so it has to be back to its
    previous value afterwards."""
    from nvidia_resiliency_ext.checkpointing.b200.persist import fast_zip_writes

    assert torch.serialization.get_crc32_options() is True
    monkeypatch.delenv("NVRX_B200_ZIP_CRC", raising=False)
    with fast_zip_writes():
        assert torch.serialization.get_crc32_options() is True
    monkeypatch.setenv("NVRX_B200_ZIP_CRC", "0")
    with fast_zip_writes():
        assert torch.serialization.get_crc32_options() is False
    assert torch.serialization.get_crc32_options() is True
    try:
        with fast_zip_writes():
            raise RuntimeError("save failed")
    except RuntimeError:
        pass
    assert torch.serialization.get_crc32_options() is True


def _w_replicated_flow(rank, world, base):
    import json

    from _replicated_manager_flow import run

    got = run(rank, os.path.join(base, "nodes"))
    with open(os.path.join(base, f"rank{rank}.json"), "w") as fh:
        json.dump(got, fh)


def test_replicated_manager_flow_matches_the_reference(tmp_path):
    """4 gloo ranks, three clique layouts: files on every node after a replicated save, find_latest and what load() returns
    after one node lost its directory (retrieve_plan + execute_plan), files after the next save (cleanup) -- all equal to what
    the imported reference did in the same flow (tests/golden/replicated_manager_4rank.json, generated by
    tests/golden/make_replicated_manager_golden.py; flow in tests/_replicated_manager_flow.py)."""
    from _replicated_manager_flow import WORLD

    run_ranks(_w_replicated_flow, WORLD, str(tmp_path), timeout=300.0)
    golden = json.load(open(GOLDEN / "replicated_manager_4rank.json"))
    for rank in range(WORLD):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        want = golden[str(rank)]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g == w, f"rank {rank}, scenario {w['scenario']}:\n mirror    {g}\n reference {w}"
            assert g["loaded_tensors"] == g["expected_tensors"]
