"""One flow through ``LocalCheckpointManager`` with clique replication on 4 gloo ranks (CPU tensors), shared by
tests/golden/make_replicated_manager_golden.py (REFERENCE package on sys.path) and tests/test_local_manager_cpu.py (mirror):
save -> what lies on every "node" -> one node loses its directory -> find_latest -> load (retrieve_plan / execute_plan) ->
second save -> cleanup.  Reference: base_manager.py:237-317,163-233, strategies.py:88-200, group_utils.py:342-497."""
import hashlib
import os
import shutil

SCENARIOS = [
    {"jump": 1, "factor": 2, "lost": 1},
    {"jump": 2, "factor": 2, "lost": 3},
    {"jump": 1, "factor": 4, "lost": 0},
]
WORLD = 4


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().view(-1).view(__import__("torch").uint8).numpy().tobytes()).hexdigest()[:16]


def rank_state(rank, iteration):
    import torch

    g = torch.Generator().manual_seed(1000 * iteration + rank)
    return {
        "w": torch.randn(17 + rank, 5, generator=g),
        "opt": [torch.randint(0, 1000, (9,), generator=g, dtype=torch.int64), {"m": torch.randn(3, generator=g).to(torch.bfloat16)}],
        "tag": f"rank{rank}@{iteration}",
        "iteration": iteration,
    }


def listing(root):
    return sorted(os.path.relpath(os.path.join(dp, f), root) for dp, _, fs in os.walk(root) for f in fs)


def run(rank, base):
    """Runs every scenario on this rank (the process group is initialised by the caller); returns this rank's record."""
    import time

    import torch.distributed as dist

    from _cpu_tasd import CpuTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    out = []
    for si, sc in enumerate(SCENARIOS):
        root = os.path.join(base, f"s{si}", f"node{rank}")
        os.makedirs(root, exist_ok=True)

        def manager():
            strat = CliqueReplicationStrategy.from_replication_params(replication_jump=sc["jump"], replication_factor=sc["factor"])
            return LocalCheckpointManager(root, repl_strategy=strat)

        rec = {"scenario": sc}
        mgr = manager()
        mgr.save(CpuTensorAwareStateDict(rank_state(rank, 9)), 9, is_async=False)
        dist.barrier()
        rec["files_after_save"] = listing(root)
        if rank == sc["lost"]:
            shutil.rmtree(root)
            os.makedirs(root)
        dist.barrier()
        mgr2 = manager()
        rec["latest"] = mgr2.find_latest()
        loaded, cid = mgr2.load()
        rec["loaded_id"] = list(cid)
        rec["loaded_tag"] = loaded.state_dict["tag"]
        rec["loaded_tensors"] = [sha(t) for t in loaded.tensors]
        rec["expected_tensors"] = [sha(t) for t in CpuTensorAwareStateDict(rank_state(rank, 9)).tensors]
        dist.barrier()
        # the next iteration: every node holds its clique's replicas again, the older iteration goes away
        mgr2.save(CpuTensorAwareStateDict(rank_state(rank, 12)), 12, is_async=False)
        dist.barrier()
        time.sleep(0.5)
        rec["files_after_second_save"] = listing(root)
        rec["latest_after_second_save"] = manager().find_latest()
        dist.barrier()
        out.append(rec)
    return out
