"""``BasicTensorAwareStateDict``: TensorAwareStateDict for plain nested dicts / lists of CUDA tensors.

API and pickle layout mirror reference ``checkpointing/local/basic_state_dict.py`` (instance attributes
``state_dict`` and ``_is_hollow``, class ``TensorPlaceholder`` with ``_device/_shape/_dtype``) so files written
by either implementation load with the other.  What differs is how the payload moves:

* ``copy_tensors_to_cpu``  (reference ``:162-174``: one ``x.to("cpu")`` per tensor) -> one pack kernel into a
  staging buffer + one side-stream drain into a pinned shared-memory slot; tensors become views of it.
* ``restore_tensor_device`` (reference ``:176-187``: one ``x.to("cuda")`` per tensor) -> one H2D copy + one
  scatter kernel.
"""

import os
from typing import Optional, Union

import torch

from .base_state_dict import TensorAwareStateDict


def nested_values(x: Union[dict, list]):
    """Depth-first iterator over the leaves of nested dicts (value order) and lists (item order).
    This traversal *is* the flattening order of the packed layout (reference ``:34-41``)."""
    children = x.values() if isinstance(x, dict) else x
    for child in children:
        if isinstance(child, (dict, list)):
            yield from nested_values(child)
        else:
            yield child


def dict_list_map_inplace(f, x):
    """Apply ``f`` to every leaf of a nested dict/list structure, rewriting containers in place."""
    if isinstance(x, dict):
        for key in x:
            x[key] = dict_list_map_inplace(f, x[key])
        return x
    if isinstance(x, list):
        x[:] = [dict_list_map_inplace(f, item) for item in x]
        return x
    return f(x)


class TensorPlaceholder:
    """What a hollow state dict keeps per tensor: device, shape and dtype (reference ``:56-72``)."""

    def __init__(self, ten):
        self._device = ten.device
        self._shape = ten.shape
        self._dtype = ten.dtype

    def init_tensor(self):
        """A new uninitialised tensor with the recorded properties."""
        return torch.empty(self._shape, dtype=self._dtype, device=self._device)


class BasicTensorAwareStateDict(TensorAwareStateDict):
    """Wraps ``state_dict`` (tensors nested only in dicts / lists, all on CUDA)."""

    def __init__(self, state_dict):
        self.state_dict = state_dict
        for leaf in nested_values(self.state_dict):
            if isinstance(leaf, torch.Tensor):
                assert leaf.is_cuda  # keeps device bookkeeping trivial: everything returns to "cuda"
        self._is_hollow = False

    # ---- payload / skeleton split -------------------------------------------------------------
    def pop_tensors(self):
        assert not self.is_hollow
        payload = list(self.tensors)
        dict_list_map_inplace(
            lambda leaf: TensorPlaceholder(leaf) if isinstance(leaf, torch.Tensor) else leaf, self.state_dict
        )
        self._is_hollow = True
        return payload

    @property
    def tensors(self):
        assert not self.is_hollow
        for leaf in nested_values(self.state_dict):
            if isinstance(leaf, torch.Tensor):
                yield leaf

    @property
    def is_hollow(self):
        return self._is_hollow

    def insert_tensors(self, tensor_data):
        assert self.is_hollow
        feed = iter(list(tensor_data))
        dict_list_map_inplace(
            lambda leaf: next(feed) if isinstance(leaf, TensorPlaceholder) else leaf, self.state_dict
        )
        self._is_hollow = False

    def init_tensors(self):
        assert self.is_hollow
        dict_list_map_inplace(
            lambda leaf: leaf.init_tensor() if isinstance(leaf, TensorPlaceholder) else leaf, self.state_dict
        )
        self._is_hollow = False

    def _replace_tensors(self, new_tensors):
        feed = iter(new_tensors)
        dict_list_map_inplace(
            lambda leaf: next(feed) if isinstance(leaf, torch.Tensor) else leaf, self.state_dict
        )

    # ---- device <-> host through the snapshot engine --------------------------------------------
    def copy_tensors_to_cpu(self, non_blocking=False, *, narrow: bool = False):
        """Snapshot all CUDA tensors into one pinned host buffer; tensors become CPU views of it.

        Returns the engine ``Snapshot`` handle (``None`` if nothing was on the GPU) so the caller can
        ``wait()`` for, follow, and finally ``release()`` the host slot; the handle is deliberately NOT kept
        on ``self`` -- the object is pickled into checkpoint files."""
        assert not self.is_hollow
        current = list(self.tensors)
        if not any(t.is_cuda for t in current):
            return None
        from ..b200.engine import SnapshotEngine

        devices = {t.device.index for t in current if t.is_cuda}
        assert len(devices) == 1, f"tensors on several devices: {devices}"
        snap = SnapshotEngine.get(devices.pop()).snapshot(current, narrow=narrow)
        self._replace_tensors(snap.host_views())
        if not non_blocking:
            snap.wait()
        return snap

    def restore_tensor_device(self, non_blocking=True, *, widen: bool = False):
        """Bring every CPU tensor to the current CUDA device (reference moves to ``"cuda"`` too, ``:184-187``).

        ``widen=True`` turns bf16 host tensors back into fp32 inside the scatter kernel (for snapshots taken
        with ``narrow=True``)."""
        assert not self.is_hollow
        current = list(self.tensors)
        host = [t for t in current if not t.is_cuda]
        if not host:
            return
        from ..b200.engine import SnapshotEngine

        widen_to: Optional[list] = None
        if widen:
            widen_to = [torch.float32 if t.dtype == torch.bfloat16 else None for t in host]
        engine = SnapshotEngine.get()
        # set by LocalCheckpointManager._load in zero-copy mode: the file may be a hard link to a slot that is still pinned in
        # this process (in-process restart), in which case the H2D reads the slot and nothing is copied on the host
        source = self.__dict__.pop("_b200_loaded_from", None)
        resident = file_source = expect = None
        if source is not None and len(host) == len(current):
            from ..b200 import fastsave, ptzip

            if os.environ.get("NVRX_B200_VERIFY_RESTORE", "0") == "1":
                # opt-in: check what arrives in HBM against the record checksums of the file (GPU kernel, see engine.restore)
                expect = ptzip.record_crcs(source, len(host))

            if fastsave.zero_copy_enabled():
                resident = engine.resident_source(source, host)
            if resident is None and os.environ.get("NVRX_B200_RESTORE_PREAD", "1") != "0":
                # default: feed the H2D straight from the file through a small pinned ring (parallel pread, chunk-pipelined
                # with the copy) instead of copying out of the file's mmap into a snapshot-sized pinned slot first
                offs = ptzip.tensor_offsets_in_file(source, host)
                file_source = (source, offs) if offs is not None else None
        moved = iter(engine.restore(host, widen_to=widen_to, resident=resident, file_source=file_source, expect_crcs=expect))
        self._replace_tensors([t if t.is_cuda else next(moved) for t in current])
