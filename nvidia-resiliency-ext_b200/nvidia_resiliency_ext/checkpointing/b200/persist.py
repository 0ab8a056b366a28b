"""Writer-process side of a snapshot: follow the drain through shared memory, then persist.

Everything in this module runs in a CPU-only process (forked child or spawned persistent worker): it never
imports the CUDA side of the engine and never calls CUDA.  The functions are module-level so that an
``AsyncRequest`` carrying them pickles for the spawn worker (the reference cannot ship
``LocalCheckpointManager._save_fn`` to its persistent worker for exactly that reason, survey appendix A.2).
"""

from __future__ import annotations

import os
from typing import Any, Dict, List

import torch

DRAIN_TIMEOUT_MS = int(os.environ.get("NVRX_B200_DRAIN_TIMEOUT_MS", str(30 * 60 * 1000)))


def fast_zip_writes() -> None:
    """Writer-process setting: do not compute zip CRC32s in ``torch.save``.

    ``torch.load`` never verifies them (checked with a zeroed CRC field, plain and ``mmap=True`` loads), but
    computing them is the dominant cost of ``torch.save`` for multi-GB payloads (single-threaded crc32 over the
    whole snapshot).  ``NVRX_B200_ZIP_CRC=1`` keeps them."""
    if os.environ.get("NVRX_B200_ZIP_CRC", "0") in ("", "0"):
        try:
            torch.serialization.set_crc32_options(False)
        except Exception:  # noqa: BLE001 - older PyTorch without the switch
            pass


def drain_aware(fn):
    """Mark ``fn`` as following the snapshot drain itself: async callers then skip the device-wide
    ``torch.cuda.synchronize()`` the reference needs before forking (``async_ckpt/core.py:345``)."""
    fn.nvrx_drain_aware = True
    return fn


class SnapshotRef:
    """Picklable stand-in for a tensor inside a state-dict skeleton: "tensor #index of the snapshot"."""

    __slots__ = ("index",)

    def __init__(self, index: int):
        self.index = index

    def __reduce__(self):
        return (SnapshotRef, (self.index,))


def _materialise(skeleton: Any, views: List[torch.Tensor]) -> Any:
    if isinstance(skeleton, dict):
        return {k: _materialise(v, views) for k, v in skeleton.items()}
    if isinstance(skeleton, list):
        return [_materialise(v, views) for v in skeleton]
    if isinstance(skeleton, SnapshotRef):
        return views[skeleton.index]
    return skeleton


def wait_for_snapshots(descs) -> list:
    """Map every snapshot slot named in ``descs`` and block (CPU only) until its drain has finished.
    Returns the mapped HostBuffers; keep them alive while tensor views are in use."""
    from .engine import HostBuffer

    held = []
    for desc in descs:
        hb = HostBuffer.open(desc["shm_name"])
        hb.wait(desc["progress_target"], DRAIN_TIMEOUT_MS)
        held.append(hb)
    return held


@drain_aware
def save_snapshot_with_torch(skeleton: Any, path, desc: Dict, *save_args, **save_kwargs) -> None:
    """``torch.save`` a state dict whose tensors live in a drained snapshot slot.

    ``skeleton`` is the user's state dict with every tensor replaced by a :class:`SnapshotRef`.  All tensors
    are views of ONE storage (the slot), so the file holds a single storage record written sequentially and
    ``torch.load`` returns tensors that compare equal to the reference's own ``torch.save`` of the CPU copies.
    """
    from . import fastsave
    from .engine import open_snapshot_views

    hb, views = open_snapshot_views(desc, DRAIN_TIMEOUT_MS)
    try:
        obj = _materialise(skeleton, views)
        fast_zip_writes()
        with fastsave.slot_ranges(fastsave.ranges_for([desc], [hb])):
            fastsave.save(obj, path, *save_args, **save_kwargs)
    finally:
        del views
        hb.close(unlink=False)
