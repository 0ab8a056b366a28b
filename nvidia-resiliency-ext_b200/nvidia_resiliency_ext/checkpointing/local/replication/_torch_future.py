"""Point-to-point object transfer (the reference back-ports ``send_object_list`` / ``recv_object_list`` from a
newer PyTorch in ``local/replication/_torch_future.py``; every PyTorch this package supports ships them)."""

from torch.distributed import recv_object_list, send_object_list  # noqa: F401
