"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Timed restatements of the reference's save / replicate / restore FLOWS.

``snapshot_oracle.py`` restates what the reference computes (bits, orders, names); this module restates HOW the
reference moves the bytes, call for call, so that ``bench.py --impl reference`` (and its ``cpu_baseline`` leg) can time
the reference's path on the GPU box, where ``/root/reference`` does not exist.  Nothing under
``nvidia-resiliency-ext_b200/`` imports it.  Paths below are relative to
``/root/reference/src/nvidia_resiliency_ext/checkpointing``.

Pinning: the flows here produce their bytes with the same PyTorch calls the reference makes (``Tensor.to``,
``torch.save`` / ``torch.load``, ``dist.broadcast``); ``tests/test_reference_port_cpu.py`` runs them on CPU tensors /
gloo and compares the files and gathered tensors with the committed outputs of the imported reference
(``tests/golden/``: ``c1_reference_async.pt``, ``iter_0000007_0_local.pt``, ``replicate_2rank.json``).
"""

from __future__ import annotations

import gc
import time
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .snapshot_oracle import flatten_tensors, map_outplace, reference_preload


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _child_save(obj, path):
    gc.disable()  # utils.py:102-120 wrap_for_async / _disable_gc
    torch.save(obj, path)


class ReferenceAsyncCheckpoint:
    """``TorchAsyncCheckpoint(persistent_queue=False)`` restated (async_ckpt/torch_ckpt.py:43-53 +
    ``TemporalAsyncCaller.schedule_async_call`` async_ckpt/core.py:318-355):

        preload_tensors (per-tensor pinned D2H on the current stream)   utils.py:85-99
        torch.cuda.synchronize()                                        torch_ckpt.py:50  <- the training stall
        torch.cuda.synchronize(); fork(); child: torch.save             core.py:345-355

    ``write=False`` skips the child (used for the bounded-sample steps of the bench: the D2H and the sync, which are what
    the stall consists of, still run on the full state)."""

    def __init__(self):
        self.process = None
        self.holder = None

    def async_save(self, state_dict, path, write: bool = True) -> None:
        self.finalize(blocking=True)  # schedule_async_request finalizes nothing itself; one call in flight keeps memory bounded
        pre = reference_preload(state_dict)
        _sync()
        self.holder = pre
        if not write:
            return
        import torch.multiprocessing as mp

        _sync()
        self.process = mp.get_context("fork").Process(target=_child_save, args=(pre, path))
        self.process.start()

    def done(self) -> bool:
        return self.process is None or not self.process.is_alive()

    def finalize(self, blocking: bool = True) -> bool:
        if self.process is not None:
            if not blocking and self.process.is_alive():
                return False
            self.process.join()
            code, self.process = self.process.exitcode, None
            if code != 0:
                raise RuntimeError(f"reference writer child failed with exit code {code}")
        self.holder = None
        return True


# ---- LocalCheckpointManager without replication: save / load of a BasicTensorAwareStateDict ---------------------------
class PortTensorAwareStateDict:
    """The part of ``BasicTensorAwareStateDict`` the save / load flow touches (local/basic_state_dict.py:162-187)."""

    def __init__(self, state_dict):
        self.state_dict = state_dict

    def copy_tensors_to_cpu(self, non_blocking=False):
        self.state_dict = map_outplace(
            lambda v: v.to("cpu", non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v, self.state_dict
        )

    def restore_tensor_device(self, non_blocking=True):
        self.state_dict = map_outplace(
            lambda v: v.to("cuda", non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v, self.state_dict
        )


def reference_local_save(state_dict, path) -> Dict[str, float]:
    """``BaseCheckpointManager.save(is_async=True)`` + ``TemporalAsyncCaller`` (local/ckpt_managers/base_manager.py:268-310,
    local_manager.py:108-130): per-tensor D2H, device sync, fork, child ``torch.save`` of the TASD object.
    Returns {"stall": s until the trainer continues, "total": s until the file is complete}."""
    import torch.multiprocessing as mp

    tasd = PortTensorAwareStateDict(state_dict)
    t0 = time.perf_counter()
    tasd.copy_tensors_to_cpu(non_blocking=True)  # base_manager.py:268
    _sync()  # base_manager.py:306-309
    _sync()  # core.py:345
    proc = mp.get_context("fork").Process(target=_child_save, args=(tasd, path))
    proc.start()
    t1 = time.perf_counter()
    proc.join()
    t2 = time.perf_counter()
    if proc.exitcode != 0:
        raise RuntimeError(f"reference writer child failed with exit code {proc.exitcode}")
    return {"stall": t1 - t0, "total": t2 - t0}


def reference_local_load(path):
    """``LocalCheckpointManager._load`` + ``_load_fn`` (local_manager.py:91-105, base_manager.py:139-145):
    ``torch.load`` then ``restore_tensor_device(non_blocking=False)`` = one blocking H2D per tensor."""
    tasd = torch.load(path, weights_only=False)
    tasd.restore_tensor_device(non_blocking=False)
    return tasd


# ---- replication ---------------------------------------------------------------------------------------------------
class _Placeholder:
    """``TensorPlaceholder`` (local/replication/torch_device_utils.py:43-99): shape / dtype travel, of the device only its TYPE --
    the receiver allocates on ITS default device of that type."""

    def __init__(self, t: torch.Tensor):
        self.shape, self.dtype, self.device_type = t.shape, t.dtype, t.device.type

    def empty_like(self):
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device_type == "cuda" else torch.device(self.device_type)
        return torch.empty(self.shape, dtype=self.dtype, device=dev)


def reference_all_gather_batch(my_tensors: Sequence[torch.Tensor], group=None, target_device: Optional[str] = "cpu") -> List[List[torch.Tensor]]:
    """``GroupWrapper.all_gather_batch`` (local/replication/group_utils.py:342-375): gather the placeholders as objects,
    then for every member, for every tensor: ``dist.broadcast`` (the member's own tensor or a fresh ``empty_like``) and
    ``.to(target_device, non_blocking=True)``.  ``CliqueReplicationStrategy.replicate`` calls it with
    ``target_device="cpu"`` (strategies.py:107-113)."""
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    gathered: List = [None] * world
    dist.all_gather_object(gathered, [_Placeholder(t) for t in my_tensors], group=group)
    out: List[List[torch.Tensor]] = []
    for r, phs in enumerate(gathered):
        src = dist.get_global_rank(group, r) if group is not None else r
        row = []
        for i, ph in enumerate(phs):
            ten = my_tensors[i] if r == me else ph.empty_like()
            dist.broadcast(ten, src=src, group=group)
            if target_device is not None:
                ten = ten.to(target_device, non_blocking=True)
            row.append(ten)
        out.append(row)
    return out


def reference_replicated_save(state_dict, paths_by_member: Sequence, group=None) -> Dict[str, float]:
    """``BaseCheckpointManager.save`` with a ``CliqueReplicationStrategy`` (base_manager.py:262-310, strategies.py:88-140):
    all-gather-batch to the host (own shard included, survey appendix A.3), sync, fork, child writes one file per member.
    ``paths_by_member[r]`` = where member r's replica goes on this rank."""
    import torch.multiprocessing as mp

    mine = flatten_tensors(state_dict)
    t0 = time.perf_counter()
    rows = reference_all_gather_batch(mine, group, "cpu")
    _sync()
    _sync()
    payload = {str(p): row for p, row in zip(paths_by_member, rows)}
    proc = mp.get_context("fork").Process(target=_child_save_many, args=(payload,))
    proc.start()
    t1 = time.perf_counter()
    proc.join()
    t2 = time.perf_counter()
    if proc.exitcode != 0:
        raise RuntimeError(f"reference writer child failed with exit code {proc.exitcode}")
    return {"stall": t1 - t0, "total": t2 - t0}


def _child_save_many(payload):
    gc.disable()
    for path, tensors in payload.items():  # base_manager.py:147-155 _save_fn loops over the ids
        torch.save(tensors, path)
