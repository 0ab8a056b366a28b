"""Host-side (metadata) cost of CliqueReplicationStrategy.replicate, measured without a GPU: F gloo ranks, the C2 structure
(1455 tensors per rank) with tiny payloads so that only pickling / object gathers / skeleton handling are timed.

    python tools/profile_replicate_metadata.py --ranks 8 --iters 5
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nvidia-resiliency-ext_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def c2_like_state(rank, numel=4):
    import torch

    g = torch.Generator().manual_seed(rank)
    names = []
    for layer in range(32):
        names += [f"layers.{layer}.{n}" for n in ("q", "k", "v", "o", "gate", "up", "down", "ln1", "ln2")]
    names += ["embed", "final_ln", "lm_head"]
    model = {n: torch.randn(numel, generator=g) for n in names}
    opt = {i: {"main_param": torch.randn(numel, generator=g), "exp_avg": torch.randn(numel, generator=g),
               "exp_avg_sq": torch.rand(numel, generator=g), "step": torch.tensor(float(rank))} for i in range(len(names))}
    return {"model": model, "optimizer": {"state": opt}, "iteration": 0}


def job(rank, world, iters, manager=False):
    import torch.distributed as dist

    from _cpu_tasd import CpuTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import CliqueReplicationStrategy

    strat = CliqueReplicationStrategy(dist.group.WORLD, target_device="cpu")
    if manager:
        import tempfile

        from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

        tmp_root = tempfile.mkdtemp(prefix="nvrx_prof_", dir="/dev/shm")
        mgr = LocalCheckpointManager(tmp_root, repl_strategy=strat)
        for it in range(iters):
            sd = c2_like_state(rank)
            sd["iteration"] = it
            tasd = CpuTensorAwareStateDict(sd)
            dist.barrier()
            prof = None
            if rank == 0 and it == iters - 1 and os.environ.get("PROFILE"):
                import cProfile

                prof = cProfile.Profile()
                prof.enable()
            t0 = time.perf_counter()
            req = mgr.save(tasd, it + 1, is_async=True)
            dt = time.perf_counter() - t0
            if prof is not None:
                import pstats

                prof.disable()
                pstats.Stats(prof).sort_stats("cumulative").print_stats(30)
            t1 = time.perf_counter()
            req.execute_sync()
            if rank == 0:
                print(f"iter {it}: manager.save() returned after {dt * 1e3:.1f} ms; write+finalize {1e3 * (time.perf_counter() - t1):.0f} ms", flush=True)
        import shutil

        dist.barrier()
        shutil.rmtree(tmp_root, ignore_errors=True)
        return
    for it in range(iters):
        sd = c2_like_state(rank)
        sd["iteration"] = it  # a changing non-tensor leaf, as a trainer would have
        tasd = CpuTensorAwareStateDict(sd)
        dist.barrier()
        prof = None
        if rank == 0 and it == iters - 1 and os.environ.get("PROFILE"):
            import cProfile

            prof = cProfile.Profile()
            prof.enable()
        t0 = time.perf_counter()
        out, ids = strat.replicate(tasd, f"id{it}_{rank}")
        dt = time.perf_counter() - t0
        if prof is not None:
            import pstats

            prof.disable()
            pstats.Stats(prof).sort_stats("cumulative").print_stats(25)
        assert len(out) == world and not out[0].is_hollow
        if rank == 0:
            print(f"iter {it}: replicate host time {dt * 1e3:.1f} ms", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--manager", action="store_true", help="time LocalCheckpointManager.save(is_async=True) instead of replicate()")
    a = ap.parse_args()
    from _mp import run_ranks

    run_ranks(job, a.ranks, a.iters, a.manager, timeout=600)
