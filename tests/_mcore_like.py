"""TEST INFRASTRUCTURE: a structural stand-in for Megatron-Core's ``MCoreTensorAwareStateDict`` (megatron.core is not
installed here; SURVEY 8(f)4, reference caller ``tests/ptl_resiliency/func/nemo20/test_local_ckpt_llama3.py:68-143``).

What it keeps of the real class, i.e. what matters to the checkpoint managers: the state is split into ``common`` (plain
Python / host tensors, pickled as is) and ``sharded_state_dict`` whose leaves are *ShardedTensor-like objects that own the
tensor in ``.data``* (plus key / global shape / offsets / replica id); ``pop_tensors`` empties ``.data`` and leaves the
sharding metadata in place; ``init_tensors`` re-creates tensors from the recorded local shape / dtype; the class brings its OWN
``copy_tensors_to_cpu`` / ``restore_tensor_device`` (per-tensor ``.to``), so an engine data path has to work through the ABC
contract alone."""
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch

from nvidia_resiliency_ext.checkpointing.local.base_state_dict import TensorAwareStateDict


@dataclass
class ShardedTensorLike:
    key: str
    data: Optional[torch.Tensor]
    dtype: torch.dtype
    local_shape: Tuple[int, ...]
    global_shape: Tuple[int, ...]
    global_offset: Tuple[int, ...]
    replica_id: Tuple[int, ...] = (0, 0, 0)
    device: str = "cuda"

    @classmethod
    def from_rank_offsets(cls, key, data, rank, world):
        shape = tuple(data.shape)
        glob = (shape[0] * world,) + shape[1:] if shape else shape
        off = (shape[0] * rank,) + (0,) * (len(shape) - 1) if shape else ()
        return cls(key, data, data.dtype, shape, glob, off, device=str(data.device.type))

    def init_data(self):
        self.data = torch.empty(self.local_shape, dtype=self.dtype, device=self.device)


@dataclass
class ShardedObjectLike:
    key: str
    data: Any


def _sharded_leaves(x):
    for v in (x.values() if isinstance(x, dict) else x):
        if isinstance(v, (dict, list)):
            yield from _sharded_leaves(v)
        elif isinstance(v, ShardedTensorLike):
            yield v


@dataclass
class MCoreLikeTensorAwareStateDict(TensorAwareStateDict):
    common: Dict[str, Any]
    sharded_state_dict: Dict[str, Any]
    _is_hollow: bool = False
    calls: list = field(default_factory=list)  # which of its own device<->host methods ran (tests look at this)

    @classmethod
    def from_state_dict(cls, model: Dict[str, torch.Tensor], optim: Dict[int, Dict[str, torch.Tensor]], rank=0, world=1, iteration=0):
        sharded = {
            "model": {k: ShardedTensorLike.from_rank_offsets(f"model.{k}", v, rank, world) for k, v in model.items()},
            "optimizer": {"state": {i: {k: ShardedTensorLike.from_rank_offsets(f"optimizer.state.{k}.{i}", v, rank, world)
                                        for k, v in st.items()} for i, st in optim.items()}},
            "rerun": ShardedObjectLike("rerun_state", {"mode": "disabled"}),
        }
        common = {"iteration": iteration, "args": {"lr": 3e-4, "tp": world}, "rng_state": torch.arange(16, dtype=torch.uint8)}
        return cls(common, sharded)

    @property
    def is_hollow(self):
        return self._is_hollow

    @property
    def tensors(self):
        assert not self._is_hollow
        return (sh.data for sh in _sharded_leaves(self.sharded_state_dict))

    def pop_tensors(self):
        assert not self._is_hollow
        out = []
        for sh in _sharded_leaves(self.sharded_state_dict):
            out.append(sh.data)
            sh.data = None
        self._is_hollow = True
        return out

    def insert_tensors(self, tensor_data):
        assert self._is_hollow
        feed = iter(list(tensor_data))
        for sh in _sharded_leaves(self.sharded_state_dict):
            sh.data = next(feed)
        self._is_hollow = False

    def init_tensors(self):
        assert self._is_hollow
        for sh in _sharded_leaves(self.sharded_state_dict):
            sh.init_data()
        self._is_hollow = False

    def copy_tensors_to_cpu(self, non_blocking=False):
        self.calls.append("copy_tensors_to_cpu")
        for sh in _sharded_leaves(self.sharded_state_dict):
            sh.data = sh.data.to("cpu", non_blocking=non_blocking)

    def restore_tensor_device(self, non_blocking=True):
        self.calls.append("restore_tensor_device")
        for sh in _sharded_leaves(self.sharded_state_dict):
            sh.data = sh.data.to("cuda", non_blocking=non_blocking)

    def to_state_dict(self):
        return {"common": self.common, "model": {k: sh.data for k, sh in self.sharded_state_dict["model"].items()}}
