"""``LocalCheckpointManager``: local checkpoints as files on node-local storage (SSD or RAM disk).

API and on-disk naming mirror reference ``checkpointing/local/ckpt_managers/local_manager.py:39-177``:
``<root>/<session_id>/<rank>/iter_<7-digit iteration>_<owner rank>_local[.dirty].pt``, written under the
``.dirty`` name with exclusive create and renamed when complete; the payload is ``torch.save`` of the
TensorAwareStateDict object, so files are interchangeable with the reference in both directions.
"""

import logging
import os
import re
from pathlib import Path
from typing import Iterable, Optional, Tuple, Union

import torch

from ...b200 import fastsave
from ...utils import debug_time
from ..base_state_dict import TensorAwareStateDict
from ..replication.strategies import ReplicationStrategy
from .base_manager import BaseCheckpointManager, CheckpointingException, CkptID, SameMachineReplicationException

logger = logging.getLogger(__name__)

_ITER_DIGITS = 7


class LocalCheckpointManager(BaseCheckpointManager):
    """File-backed local checkpoint manager.

    Args:
        root_local_ckpt_dir: root directory on local storage; iterations share it (names are unique).
        session_id: extra path component to separate workloads sharing an administrator-chosen root.
        repl_strategy: optional replication of shards to other ranks.
    """

    def __init__(
        self,
        root_local_ckpt_dir: Union[str, Path],
        session_id: str = "",
        repl_strategy: Optional[ReplicationStrategy] = None,
    ):
        super().__init__(session_id, repl_strategy)
        self.root_local_ckpt_dir = root_local_ckpt_dir
        self._dir_created = False
        self._local_ckpt_dir = None

    @property
    def local_ckpt_dir(self):
        if self._local_ckpt_dir is None:
            self._local_ckpt_dir = Path(self.root_local_ckpt_dir) / self.session_id / str(self.rank)
        return self._local_ckpt_dir

    def _ensure_dir(self):
        if not self._dir_created:
            os.makedirs(self.local_ckpt_dir, exist_ok=True)
            self._dir_created = True

    # ---- naming -------------------------------------------------------------------------------
    def _filename_from_template(self, iteration: Union[int, str], rank: Union[int, str], extra_suffix: str = ""):
        """``iter_<iteration>_<rank>_local<extra_suffix>.pt``; integer iterations are zero-padded to 7 digits,
        strings (glob / regex fragments) are used verbatim."""
        it = str(iteration).zfill(_ITER_DIGITS) if isinstance(iteration, int) else iteration
        if it.isdigit():
            assert len(it) == _ITER_DIGITS
        return f"iter_{it}_{rank}_local{extra_suffix}.pt"

    def _local_ckpt_path_from_id(self, ckpt_id, is_dirty=False):
        iteration, rank, session_id = ckpt_id
        assert session_id == self.session_id
        return self.local_ckpt_dir / self._filename_from_template(iteration, rank, ".dirty" if is_dirty else "")

    def _filename_to_id(self, filename):
        _, iteration, rank, _ = filename.split("_", 3)
        return (int(iteration), int(rank), self.session_id)

    # ---- backend hooks ------------------------------------------------------------------------
    def _my_ckpt_ids(self) -> Iterable[CkptID]:
        self._ensure_dir()
        # '\\' as suffix escapes the dot: iter_\d+_\d+_local\.pt  (".dirty" files never match)
        complete = re.compile(self._filename_from_template("\\d+", "\\d+", "\\"))
        return [
            self._filename_to_id(entry.name)
            for entry in self.local_ckpt_dir.iterdir()
            if entry.is_file() and complete.fullmatch(entry.name)
        ]

    @debug_time("LocalCheckpointManager._load", logger)
    def _load(self, ckpt_id: CkptID) -> Tuple[TensorAwareStateDict, str]:
        path = self._local_ckpt_path_from_id(ckpt_id)
        try:
            try:
                # map the file instead of copying it: tensors are read once more anyway (parallel gather into the
                # pinned slot, then one H2D + scatter kernel)
                loaded = torch.load(path, weights_only=False, mmap=True)  # nosec B614 - files are produced by this manager
                if hasattr(loaded, "__dict__"):
                    loaded.__dict__["_b200_loaded_from"] = str(path)  # consumed by restore_tensor_device
                return loaded
            except (RuntimeError, ValueError):
                return torch.load(path, weights_only=False)  # nosec B614 - legacy (non-zip) or unmappable file
        except FileNotFoundError as exc:
            msg = f"File {path} does not exist!"
            logging.info(msg)
            logger.debug(f"{msg}. Checkpoint directory content: {[f.name for f in self.local_ckpt_dir.iterdir()]}")
            raise CheckpointingException(msg) from exc

    @debug_time("LocalCheckpointManager._save", logger)
    def _save(self, state_dict: TensorAwareStateDict, ckpt_id: CkptID):
        self._ensure_dir()
        dirty = self._local_ckpt_path_from_id(ckpt_id, True)
        assert ".dirty" in dirty.suffixes
        try:
            logging.info(f"Saving to {str(dirty)}")
            with open(dirty, "x+b") as fh:  # exclusive create: a second writer on this machine must fail
                fastsave.save(state_dict, fh)  # torch.save format; payload by parallel pwrite when it sits in a slot  # nosec B614
            final = self._local_ckpt_path_from_id(ckpt_id, False)
            logging.info(f"Renaming {str(dirty)} to {final}")
            dirty.rename(target=final)
        except FileExistsError as exc:
            logger.debug(f"Checkpoint directory content: {[f.name for f in self.local_ckpt_dir.iterdir()]}")
            raise SameMachineReplicationException(ckpt_id) from exc

    @debug_time("LocalCheckpointManager._cleanup", logger)
    def _cleanup(self, iteration):
        """Delete every checkpoint file (dirty ones included) older than ``iteration``."""
        for path in list(self.local_ckpt_dir.glob(self._filename_from_template("*", "*", "*"))):
            if self._filename_to_id(path.name)[0] < iteration:
                logging.info(f"Removing {path}")
                path.unlink()

    @debug_time("LocalCheckpointManager._cleanup_failed_save", logger)
    def _cleanup_failed_save(self, iteration):
        """Delete whatever a failed save of ``iteration`` left behind."""
        for path in list(self.local_ckpt_dir.glob(self._filename_from_template(iteration, "*", "*"))):
            logging.info(f"Removing {path}")
            path.unlink()
