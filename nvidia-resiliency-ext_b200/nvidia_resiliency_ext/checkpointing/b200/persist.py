"""Writer-process side of a snapshot: follow the drain through shared memory, then persist.

Everything in this module runs in a CPU-only process (forked child or spawned persistent worker): it never
imports the CUDA side of the engine and never calls CUDA.  The functions are module-level so that an
``AsyncRequest`` carrying them pickles for the spawn worker (the reference cannot ship
``LocalCheckpointManager._save_fn`` to its persistent worker for exactly that reason, survey appendix A.2).
"""

from __future__ import annotations

import os
from contextlib import contextmanager
from typing import Any, Dict, List

import torch

DRAIN_TIMEOUT_MS = int(os.environ.get("NVRX_B200_DRAIN_TIMEOUT_MS", str(30 * 60 * 1000)))


def zip_crc_enabled() -> bool:
    """Record checksums of checkpoint containers are ON by default, as in the files the reference's ``torch.save`` writes
    (round 1 defaulted to off).  ``NVRX_B200_ZIP_CRC=0`` leaves the data records' CRC fields zero: ``torch.load`` never
    verifies them, ``zipfile.testzip()`` and other zip tools would complain."""
    return os.environ.get("NVRX_B200_ZIP_CRC", "1") not in ("", "0")


@contextmanager
def fast_zip_writes():
    """While saving a snapshot with ``NVRX_B200_ZIP_CRC=0``: do not compute zip CRC32s in ``torch.save`` either.

    By default nothing is switched: the payload written from snapshot slots is summed by the slot's thread pool
    (``nvrx_hostbuf_crc32v``, carry-less-multiply folding, ~6 GB/s per thread instead of ~1 GB/s on the one thread of
    PyTorch's writer) and patched into the container (``ptzip.patch_record_crcs``); whatever goes through stock ``torch.save``
    keeps PyTorch's own checksums.  The process-wide PyTorch switch is put back afterwards: a synchronous save runs in the
    trainer, whose own ``torch.save`` calls must keep their checksums."""
    previous = None
    if not zip_crc_enabled():
        try:
            previous = torch.serialization.get_crc32_options()
            torch.serialization.set_crc32_options(False)
        except Exception:  # noqa: BLE001 - older PyTorch without the switch
            previous = None
    try:
        yield
    finally:
        if previous is not None:
            torch.serialization.set_crc32_options(previous)


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


# Slot mappings kept by a long-lived writer (the persistent worker): re-mapping a 16 GB slot for every checkpoint costs
# millions of page faults and a page-table teardown per save; slots are few and reused by the engine.
_slot_cache: Dict[str, Any] = {}
_SLOT_CACHE_MAX = 8


class _Borrowed:
    """A cached HostBuffer handed to code that calls ``close()`` when done: closing a borrowed mapping is a no-op."""

    def __init__(self, hb):
        self._hb = hb

    def __getattr__(self, name):
        return getattr(self._hb, name)

    def close(self, unlink=None):
        return None


def open_slot(name: str, cache: bool):
    from .engine import HostBuffer

    if not cache:
        return HostBuffer.open(name)
    hb = _slot_cache.get(name)
    if hb is None:
        for stale in [n for n in _slot_cache if not os.path.exists("/dev/shm" + n)]:
            _slot_cache.pop(stale).close(unlink=False)
        while len(_slot_cache) >= _SLOT_CACHE_MAX:
            _slot_cache.pop(next(iter(_slot_cache))).close(unlink=False)
        hb = _slot_cache[name] = HostBuffer.open(name)
    return _Borrowed(hb)


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
    from .engine import host_views

    # the persistent worker (it sets NVRX_B200_CACHE_SLOTS in its loop) keeps slots mapped between checkpoints; a forked
    # one-shot child or an inline call maps and unmaps
    hb = open_slot(desc["shm_name"], cache=os.environ.get("NVRX_B200_CACHE_SLOTS") == "1")
    hb.wait(desc["progress_target"], DRAIN_TIMEOUT_MS)
    views = host_views(desc["layout"], hb)
    try:
        obj = _materialise(skeleton, views)
        with fast_zip_writes(), fastsave.slot_ranges(fastsave.ranges_for([desc], [hb])):
            fastsave.save(obj, path, *save_args, **save_kwargs)
    finally:
        del views
        hb.close(unlink=False)
