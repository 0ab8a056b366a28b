"""The ``TensorAwareStateDict`` contract between user state dicts and checkpoint managers.

Mirror of reference ``checkpointing/local/base_state_dict.py:29-120`` (same abstract members, same
semantics).  A TensorAwareStateDict separates *tensor payload* from *picklable skeleton*: the manager can
pop the tensors (leaving a "hollow" skeleton that pickles in microseconds), move / pack / exchange the
payload with bulk primitives, and insert it back.  The order in which ``tensors`` yields is the
**flattening order**; it defines the segment order of the packed staging buffer.
"""

from abc import ABC, abstractmethod
from typing import Any, Iterable, Sequence, ValuesView

import torch


class TensorAwareStateDict(ABC):
    """Interface a state dict must offer so a checkpoint manager can migrate its tensors efficiently."""

    @abstractmethod
    def pop_tensors(self) -> Sequence[torch.Tensor]:
        """Detach the tensor payload and return it (flattening order).

        The skeleton keeps what is needed to re-create empty tensors (shape, dtype, device) and becomes
        *hollow*; popping a hollow state dict is an error."""

    @property
    @abstractmethod
    def tensors(self) -> Iterable[torch.Tensor]:
        """The tensor payload in flattening order (state dict must not be hollow)."""

    @property
    @abstractmethod
    def is_hollow(self) -> bool:
        """True between ``pop_tensors`` and the next ``insert_tensors`` / ``init_tensors``."""

    @abstractmethod
    def insert_tensors(self, tensor_data: Iterable[torch.Tensor]):
        """Inverse of ``pop_tensors``: ``sd.insert_tensors(sd.pop_tensors())`` leaves ``sd`` unchanged."""

    @abstractmethod
    def init_tensors(self):
        """Fill a hollow state dict with freshly allocated, uninitialised tensors of the recorded
        shape / dtype / device."""

    @abstractmethod
    def copy_tensors_to_cpu(self, non_blocking=False):
        """Replace every tensor by a host copy (the device originals are not destroyed).  With
        ``non_blocking=True`` the copies are only valid once the transfer has been waited for."""

    @abstractmethod
    def restore_tensor_device(self, non_blocking=True):
        """Move every tensor back to its compute device if it is not there already."""

    def values(self) -> ValuesView[Any]:
        """Values of the instance dictionary (what the manager walks for non-tensor content)."""
        return vars(self).values()
