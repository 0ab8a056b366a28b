"""A TensorAwareStateDict for CPU-only boxes: BasicTensorAwareStateDict minus the device moves, so the manager /
replication host logic can be exercised under gloo without a GPU."""
from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict


class CpuTensorAwareStateDict(BasicTensorAwareStateDict):
    def __init__(self, state_dict):
        self.state_dict = state_dict
        self._is_hollow = False

    def copy_tensors_to_cpu(self, non_blocking=False):
        return None

    def restore_tensor_device(self, non_blocking=True):
        return None
