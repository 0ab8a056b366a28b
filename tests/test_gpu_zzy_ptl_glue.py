"""Row a16 on a GPU: the PyTorch-Lightning glue driving the real LocalCheckpointManager on the engine.  In a file of its own,
late in the alphabet: it was written after the round's last GPU visit (its CPU twin runs in tests/test_engine_flow_cpu.py), and
with ``-x`` a surprise here must not hide the parity tests."""
import pytest
import torch

from test_gpu_api import bit_equal, model_state

pytestmark = pytest.mark.gpu

from test_ptl_glue_cpu import glue  # noqa: E402,F401  (fixture: minimal stub of the three lightning symbols the glue imports)


def test_ptl_glue_drives_the_local_manager_on_gpu(glue, shm_dir, dist_1rank, built_library):  # noqa: F811
    """Row a16: ``LocalCheckpointCallback`` -> ``trainer.save_checkpoint(None, storage_options=...)`` ->
    ``HierarchicalCheckpointIO`` -> the REAL ``LocalCheckpointManager`` on the engine -> ``AsyncCallsQueue`` -> resume from the
    local checkpoint (reference ptl_resiliency/local_checkpoint_callback.py:53-212; lightning itself is not installed, its
    three symbols are stubbed)."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    class IO(glue.HierarchicalCheckpointIO):
        def to_tensor_aware_state_dict(self, checkpoint):
            return BasicTensorAwareStateDict(checkpoint)

        def from_tensor_aware_state_dict(self, tasd, **kw):
            return tasd.state_dict

    class GlobalIO:
        def load_checkpoint(self, path, map_location=None, **kw):
            return {"from": "global"}

        def save_checkpoint(self, *a, **k):
            raise AssertionError("a local save must not reach the global CheckpointIO")

    mgr = LocalCheckpointManager(shm_dir / "ptl")
    io = IO(GlobalIO(), mgr, get_global_ckpt_iteration_fn=lambda p: int(str(p).rsplit("=", 1)[-1]), async_save=True)
    q = AsyncCallsQueue(persistent=False)
    state = model_state(ntensor=6, size=(513, 255), seed=21)
    want = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in state.items() if k.startswith("param_")}

    class Trainer:
        global_step = 40

        def save_checkpoint(self, path, storage_options=None):
            req = io.save_checkpoint({"state_dict": {k: state[k] for k in want}, "global_step": self.global_step}, path, storage_options)
            q.schedule_async_request(req)

    try:
        cb = glue.LocalCheckpointCallback(every_n_train_steps=20)
        cb._save_last_checkpoint(Trainer(), {})
        for v in state.values():
            if isinstance(v, torch.Tensor) and v.is_floating_point():
                v.zero_()  # training goes on
        q.maybe_finalize_async_calls(blocking=True, no_dist=False)
        assert io.load_checkpoint("/global/step=50") == {"from": "global"}  # the global one is newer
        io2 = IO(GlobalIO(), LocalCheckpointManager(shm_dir / "ptl"), get_global_ckpt_iteration_fn=lambda p: 30)
        resumed = io2.load_checkpoint("/global/step=30")  # local (40) is newer
        assert resumed["global_step"] == 40
        assert all(resumed["state_dict"][k].is_cuda and bit_equal(resumed["state_dict"][k], w) for k, w in want.items())
    finally:
        q.close()
