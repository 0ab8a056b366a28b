"""Lightning entry points into local checkpointing: ``LocalCheckpointCallback`` + ``HierarchicalCheckpointIO``.

API mirror of reference ``ptl_resiliency/local_checkpoint_callback.py:53-212`` -- the caller glue NeMo uses to reach
``LocalCheckpointManager.save / find_latest / load``.  Nothing here moves tensor data; it decides
*which* checkpoint path (local manager vs. the wrapped global ``CheckpointIO``) a save or load takes.  Needs
``lightning`` (or ``pytorch_lightning``) at import time, like the reference.
"""

import importlib.util
import logging
from abc import abstractmethod
from datetime import timedelta
from functools import partial
from typing import Any, Callable, Dict, NewType, Optional

from ..checkpointing.async_ckpt.core import AsyncRequest
from ..checkpointing.local.base_state_dict import TensorAwareStateDict
from ..checkpointing.local.ckpt_managers.base_manager import BaseCheckpointManager

if importlib.util.find_spec("lightning") is not None:
    import lightning.pytorch as pl
    from lightning.pytorch.plugins.io.wrapper import _WrappingCheckpointIO
elif importlib.util.find_spec("pytorch_lightning") is not None:
    import pytorch_lightning as pl
    from pytorch_lightning.plugins.io.wrapper import _WrappingCheckpointIO
else:
    raise ImportError("Could not find 'lightning' or 'pytorch_lightning' module")

logger = logging.getLogger(__name__)

StateDict = NewType("StateDict", Any)
LOCAL_CKPT_OPTS_KEY = "local_checkpoint_options"


class LocalCheckpointCallback(pl.callbacks.ModelCheckpoint):
    """``ModelCheckpoint`` reduced to "save a local checkpoint every N train steps / T wall time".

    Epoch-end, validation-end and top-k saves are disabled (local checkpoints are ephemeral); the periodic
    "last" save is turned into ``trainer.save_checkpoint(None, storage_options={LOCAL_CKPT_OPTS_KEY: ...})`` which
    ``HierarchicalCheckpointIO`` recognises.  Must be used together with ``HierarchicalCheckpointIO``."""

    def __init__(self, every_n_train_steps: Optional[int] = None, train_time_interval: Optional[timedelta] = None):
        super().__init__(every_n_train_steps=every_n_train_steps, train_time_interval=train_time_interval)

    def on_train_epoch_end(self, trainer, pl_module) -> None:
        logger.info("Skipping on_train_epoch_end local ckpt save")

    def on_validation_end(self, trainer, pl_module) -> None:
        logger.info("Skipping on_validation_end local ckpt save")

    def _save_topk_checkpoint(self, trainer, monitor_candidates) -> None:
        logger.info("Skipping _save_topk_checkpoint local ckpt save")

    def _save_last_checkpoint(self, trainer, monitor_candidates) -> None:
        opts = dict(ckpt_type="local", iteration=trainer.global_step)
        trainer.save_checkpoint(None, storage_options={LOCAL_CKPT_OPTS_KEY: opts})


class HierarchicalCheckpointIO(_WrappingCheckpointIO):
    """Wraps a global ``CheckpointIO``; saves carrying ``LOCAL_CKPT_OPTS_KEY`` go to ``local_ckpt_manager``, loads
    resume from whichever of the local / global checkpoints is newer.

    Args:
        wrapped_checkpoint_io: the global CheckpointIO.
        local_ckpt_manager: manager used for local checkpoints.
        get_global_ckpt_iteration_fn: maps a global checkpoint path to its iteration.
        async_save: default ``is_async`` for local saves (overridable per save through the options).
    Subclasses provide the conversion between Lightning checkpoints and ``TensorAwareStateDict``."""

    def __init__(
        self,
        wrapped_checkpoint_io,
        local_ckpt_manager: BaseCheckpointManager,
        get_global_ckpt_iteration_fn: Callable[[Any], int],
        async_save: bool = False,
    ):
        super().__init__(wrapped_checkpoint_io)
        self.local_ckpt_manager = local_ckpt_manager
        self.get_global_ckpt_iteration_fn = get_global_ckpt_iteration_fn
        self.async_save = async_save

    def save_checkpoint(self, checkpoint: Dict[str, Any], path, storage_options: Optional[Any] = None) -> Optional[AsyncRequest]:
        if storage_options is None or LOCAL_CKPT_OPTS_KEY not in storage_options:
            return self.checkpoint_io.save_checkpoint(checkpoint, path, storage_options)
        if path is not None:
            raise ValueError(f"Path shouldn't be set for a local checkpoint, got: {path}.")
        return self._save_local_checkpoint(checkpoint, storage_options.get(LOCAL_CKPT_OPTS_KEY))

    def _save_local_checkpoint(self, checkpoint: Dict[str, Any], local_ckpt_options: dict) -> Optional[AsyncRequest]:
        return self.local_ckpt_manager.save(
            self.to_tensor_aware_state_dict(checkpoint),
            local_ckpt_options["iteration"],
            is_async=local_ckpt_options.get("is_async", self.async_save),
        )

    def load_checkpoint(self, path, map_location: Optional[Any] = None, **kwargs) -> Dict[str, Any]:
        local_it = self.local_ckpt_manager.find_latest()
        if local_it < 0:
            logger.debug("No local checkpoint available")
            return self.checkpoint_io.load_checkpoint(path, map_location=map_location, **kwargs)
        global_it = self.get_global_ckpt_iteration_fn(path)
        if local_it >= global_it:
            logger.info(
                f"Local checkpoint interation {local_it} greater than global {global_it}. Resuming from a local checkpoint"
            )
            tasd, name = self.local_ckpt_manager.load()
            logger.debug(f"Loaded local checkpoint {name}")
            return self.from_tensor_aware_state_dict(tasd, **kwargs)
        logger.warning(
            f"Found available local checkpoint from interation {local_it}, but global iteration {global_it} is greater."
            f" Resuming from a global checkpoint."
        )
        return self.checkpoint_io.load_checkpoint(path, map_location=map_location, **kwargs)

    def remove_checkpoint(self, path) -> None:
        """Local checkpoints are cleaned up by the manager itself; only global ones are removed here."""
        return self.checkpoint_io.remove_checkpoint(path)

    @classmethod
    def get_partial_wrapper_constructor(cls, local_ckpt_manager: BaseCheckpointManager, get_global_ckpt_iteration_fn: Callable[[Any], int]):
        """Constructor with everything bound except the wrapped CheckpointIO."""
        return partial(cls, local_ckpt_manager=local_ckpt_manager, get_global_ckpt_iteration_fn=get_global_ckpt_iteration_fn)

    @abstractmethod
    def to_tensor_aware_state_dict(self, checkpoint: Dict[str, Any]) -> TensorAwareStateDict:
        raise NotImplementedError

    @abstractmethod
    def from_tensor_aware_state_dict(self, tensor_aware_checkpoint: TensorAwareStateDict, **kwargs):
        raise NotImplementedError
