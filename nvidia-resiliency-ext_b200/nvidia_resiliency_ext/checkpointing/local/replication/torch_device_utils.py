"""Device-agnostic tensor placeholders used when exchanging tensor metadata between ranks
(mirror of reference ``local/replication/torch_device_utils.py:19-99``)."""

import torch


def get_default_device_from_type(device_type: str) -> torch.device:
    """``"cpu"`` -> cpu, ``"cuda"`` -> the current CUDA device; anything else is an error."""
    if device_type == "cpu":
        return torch.device("cpu")
    if device_type == "cuda":
        return torch.device(f"cuda:{torch.cuda.current_device()}")
    raise ValueError(f"Device type {device_type} unsupported!")


class TensorPlaceholder:
    """Shape/dtype/stride of a tensor (kept as a ``meta`` tensor) plus the *type* of device it lived on.
    Picklable and free of payload; a receiver can allocate a matching buffer on its own device."""

    def __init__(self, tensor: torch.Tensor):
        self.hollow_tensor = torch.empty_like(tensor, device="meta")
        self.orig_device_type = tensor.device.type
        self._nbytes = tensor.numel() * tensor.element_size()  # asked for several times per exchange and placeholder

    @property
    def device(self):
        return get_default_device_from_type(self.orig_device_type)

    @property
    def nbytes(self) -> int:
        cached = self.__dict__.get("_nbytes")  # absent on placeholders unpickled from an older peer
        return cached if cached is not None else self.hollow_tensor.numel() * self.hollow_tensor.element_size()

    def empty_like(self, device=None):
        """Uninitialised tensor of the recorded shape on ``device`` (default: local device of the original type)."""
        return torch.empty_like(self.hollow_tensor, device=self.device if device is None else device)

    def restore(self, data=None) -> torch.Tensor:
        """``data`` moved to the local device of the original type (checked against the recorded shape and
        dtype), or an empty tensor when ``data`` is None."""
        if data is None:
            return torch.empty_like(self.hollow_tensor, device=self.device)
        assert self.hollow_tensor.shape == data.shape
        assert self.hollow_tensor.dtype == data.dtype
        return data.to(self.device)
