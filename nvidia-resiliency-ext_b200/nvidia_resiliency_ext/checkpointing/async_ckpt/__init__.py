"""Asynchronous ``torch.save`` (``TorchAsyncCheckpoint``) and the async-call queue behind it."""

from .core import AsyncCallsQueue, AsyncRequest, abort_nvrx_checkpoint  # noqa: F401
from .torch_ckpt import TorchAsyncCheckpoint  # noqa: F401
