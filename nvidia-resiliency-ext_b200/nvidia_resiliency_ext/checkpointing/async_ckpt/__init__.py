"""Asynchronous ``torch.save`` (``TorchAsyncCheckpoint``), the async-call queue behind it, and the async writer for
``torch.distributed.checkpoint`` (``filesystem_async.FileSystemWriterAsync`` + ``state_dict_saver``)."""

from .core import AsyncCallsQueue, AsyncRequest, abort_nvrx_checkpoint  # noqa: F401
from .torch_ckpt import TorchAsyncCheckpoint  # noqa: F401
