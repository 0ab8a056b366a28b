"""``BaseCheckpointManager``: save / find_latest / load orchestration for local checkpoints.

API mirror of reference ``checkpointing/local/ckpt_managers/base_manager.py:64-317`` (same hooks for
subclasses, same exceptions, same collective contract: every method is called by all ranks in the same
order).  Differences, all on the snapshot path:

* ``save`` does not end in ``torch.cuda.synchronize()`` (reference ``:306-309``).  For state dicts that
  implement the full TensorAwareStateDict contract the payload goes through the snapshot engine (one pack
  kernel, one side-stream drain) and the returned ``AsyncRequest`` is *drain aware*: its writer follows the
  drain through shared memory.  State dicts that only implement ``copy_tensors_to_cpu`` keep the reference
  behaviour (host copies + sync before the request is returned).
* ``load`` restores with one H2D copy + one scatter kernel (``_load_fn`` -> ``restore_tensor_device``).

Set ``NVRX_B200_EAGER_SYNC=1`` to wait for the drain before ``save`` returns (reference timing semantics).
"""

import logging
import os
from abc import ABC, abstractmethod
from contextlib import nullcontext
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Iterable, Optional, Tuple

import torch

from ...async_ckpt.core import AsyncRequest
from ...b200 import fastsave
from ...b200.persist import fast_zip_writes, wait_for_snapshots
from ...utils import _disable_gc, debug_time
from ..base_state_dict import TensorAwareStateDict
from ..replication.group_utils import GroupWrapper
from ..replication.strategies import ReplicationStrategy

logger = logging.getLogger(__name__)

CkptID = Tuple[int, int, Any]  # (iteration, owner rank, session id)


class CheckpointingException(Exception):
    """Base class of local-checkpointing errors."""


class SameMachineReplicationException(CheckpointingException):
    """A replica would overwrite a checkpoint file that already exists on this machine."""

    def __init__(self, ckpt_id):
        super().__init__(f"Checkpoint '{ckpt_id}' already exists on the same machine.")


def _eager_sync() -> bool:
    return os.environ.get("NVRX_B200_EAGER_SYNC", "0") not in ("", "0", "false", "False")


class BaseCheckpointManager(ABC):
    """Backend-independent part of a local checkpoint manager (replication hooks included)."""

    def __init__(self, session_id, repl_strategy: ReplicationStrategy = None):
        self.latest_iteration = -1
        self.repl_strategy = repl_strategy
        self.session_id = session_id
        self._rank = None
        self._outstanding = []  # (weak reference to the save's finalize_fn, [Snapshot]) of saves not finalized yet

    @property
    def rank(self):
        if self._rank is None:
            if torch.distributed.is_initialized():
                self._rank = torch.distributed.get_rank()
            else:
                logger.warning("Torch distributed backend has not been initialized.")
                self._rank = 0
        return self._rank

    def _ckpt_id(self, iteration: int) -> CkptID:
        """This rank's checkpoint id for ``iteration``."""
        if iteration < 0:
            raise CheckpointingException(f"Invalid iteration: expected a non-negative value, got {iteration}.")
        return (iteration, self.rank, self.session_id)

    # ---- backend hooks ------------------------------------------------------------------------
    @abstractmethod
    def _my_ckpt_ids(self) -> Iterable[CkptID]:
        """Ids of the checkpoints stored on this rank."""

    @abstractmethod
    def _load(self, ckpt_id: CkptID) -> TensorAwareStateDict:
        """Read one checkpoint; raise ``CheckpointingException`` on failure."""

    @abstractmethod
    def _save(self, state_dict: TensorAwareStateDict, ckpt_id: CkptID):
        """Write one checkpoint; raise ``SameMachineReplicationException`` if it already exists."""

    @abstractmethod
    def _cleanup(self, iteration):
        """Remove what became obsolete once ``iteration`` was saved successfully."""

    @abstractmethod
    def _cleanup_failed_save(self, iteration):
        """Remove the partial results of a failed save of ``iteration``."""

    # ---- load ---------------------------------------------------------------------------------
    @debug_time("BaseCheckpointManager._load_fn", logger)
    def _load_fn(self, ckpt_id: CkptID) -> TensorAwareStateDict:
        state_dict = self._load(ckpt_id)
        if not self._restore_through_engine(state_dict):
            state_dict.restore_tensor_device(non_blocking=False)
        logger.debug(f"Finish loading {ckpt_id}")
        return state_dict

    def _restore_through_engine(self, state_dict) -> bool:
        """Mirror of :meth:`_snapshot_state_dict` for third-party state dicts that implement the whole contract: pop the host
        tensors, restore them with the engine (file -> pinned ring -> device when the backend noted the file, else one H2D;
        one scatter kernel), insert the device tensors back.  ``BasicTensorAwareStateDict`` does this itself."""
        from ..basic_state_dict import BasicTensorAwareStateDict

        if isinstance(state_dict, BasicTensorAwareStateDict) or os.environ.get("NVRX_B200_GENERIC_TASD", "1") == "0":
            return False
        if not torch.cuda.is_available():
            return False
        try:
            if state_dict.is_hollow:
                return False
            host = list(state_dict.pop_tensors())
        except NotImplementedError:
            return False
        if not host or any(t.is_cuda for t in host):
            state_dict.insert_tensors(host)
            return False
        from ...b200 import ptzip
        from ...b200.engine import SnapshotEngine

        source = getattr(state_dict, "__dict__", {}).pop("_b200_loaded_from", None)
        file_source = None
        if source is not None and os.environ.get("NVRX_B200_RESTORE_PREAD", "1") != "0":
            offs = ptzip.tensor_offsets_in_file(source, host)
            file_source = (source, offs) if offs is not None else None
        state_dict.insert_tensors(SnapshotEngine.get().restore(host, file_source=file_source))
        return True

    # ---- save (runs in the writer process) ------------------------------------------------------
    @debug_time("BaseCheckpointManager._save_fn", logger)
    @_disable_gc()
    def _save_fn(self, id_to_state_dict, snapshot_descs=()):
        held = wait_for_snapshots(snapshot_descs)  # CPU-only wait for the drain(s); no CUDA in the writer
        ckpt_id = None
        try:
            # backends that save through b200.fastsave.save() get the payload written in parallel from the slots
            with (fast_zip_writes() if snapshot_descs else nullcontext()), fastsave.slot_ranges(fastsave.ranges_for(snapshot_descs, held)):
                for ckpt_id, state_dict in id_to_state_dict.items():
                    try:
                        self._save(state_dict, ckpt_id)
                    except Exception as exc:
                        logging.error(f"Exception caught during saving {ckpt_id}: {exc}", exc_info=True)
                        raise
        finally:
            for hb in held:
                hb.close(unlink=False)
        logger.debug(f"Finish saving {ckpt_id}")

    # the bound method is what async callers run; mark it as following the drain itself
    _save_fn.nvrx_drain_aware = True

    @debug_time("BaseCheckpointManager.find_latest", logger)
    def find_latest(self):
        """Newest iteration for which *every* rank's shard is available somewhere (-1 if none).  Collective.

        The result is cached until the next ``save``."""
        if self.latest_iteration != -1:
            logger.debug(f"Using cached latest_iteration: {self.latest_iteration} in find_latest")
            return self.latest_iteration
        group = GroupWrapper()
        mine = self._my_ckpt_ids()
        if self.repl_strategy is None:
            # without replication nobody can serve another rank's shard: foreign files do not count
            mine = [cid for cid in mine if cid[1] == self.rank]
        self.globally_available_ids = group.all_gather_object(mine)

        covered = defaultdict(set)
        for ids in self.globally_available_ids:
            for iteration, owner, session in ids:
                assert type(iteration) is int
                assert session == self.session_id
                covered[iteration].add(owner)
        everyone = set(group.ranks)
        self.latest_iteration = max((it for it, owners in covered.items() if owners == everyone), default=-1)
        return self.latest_iteration

    @debug_time("BaseCheckpointManager.load", logger)
    def load(self) -> Tuple[TensorAwareStateDict, str]:
        """Load the checkpoint ``find_latest`` selected.  Collective.  Returns ``(state_dict, ckpt_id)`` with
        tensors on the compute device."""
        if self.latest_iteration == -1:
            raise CheckpointingException("The 'find_latest' method must be called before invoking the 'load' function.")
        ckpt_id = self._ckpt_id(self.latest_iteration)
        logger.debug(f"Loading checkpoint from {self.latest_iteration} iteration")
        if self.repl_strategy is None:
            return self._load_fn(ckpt_id), ckpt_id
        plan = self.repl_strategy.retrieve_plan(self.globally_available_ids, [ckpt_id])
        to_send = {cid: self._load_fn(cid) for cid in plan.required_ids()}
        received = list(self.repl_strategy.retrieve_execute(plan, to_send).items())
        assert len(received) == 1, f"Got {len(received)} IDs, but requested only 1!"
        assert received[0][0] == ckpt_id, f"Retrieved different ID ({received[0][0]}) than requested ({ckpt_id})?"
        return received[0][1], ckpt_id

    # ---- save (trainer side) ----------------------------------------------------------------------
    def _snapshot_state_dict(self, state_dict: TensorAwareStateDict):
        """Move the payload to the host.  Returns the engine Snapshot, or None when the state dict made its own
        (reference-style) host copies, in which case a device sync is required before the writer may read them.

        * ``BasicTensorAwareStateDict`` snapshots through the engine itself.
        * Any OTHER TensorAwareStateDict that implements the whole contract (``pop_tensors`` / ``insert_tensors``, e.g.
          Megatron-Core's ``MCoreTensorAwareStateDict``, whose tensors sit inside ShardedTensor objects) gets the same
          data path through the contract alone: pop the payload, ONE pack + drain, insert the host views back.  That is the
          round trip ``CliqueReplicationStrategy.replicate`` already relies on (reference ``strategies.py:100-131``).
        * A state dict that only implements ``copy_tensors_to_cpu`` (the reference's ``SimpleTensorAwareStateDict`` test
          class raises ``NotImplementedError`` everywhere else) keeps the reference behaviour.
        ``NVRX_B200_GENERIC_TASD=0`` turns the second case off."""
        from ..basic_state_dict import BasicTensorAwareStateDict

        if isinstance(state_dict, BasicTensorAwareStateDict):
            result = state_dict.copy_tensors_to_cpu(non_blocking=True)
            return result if hasattr(result, "descriptor") else None
        if os.environ.get("NVRX_B200_GENERIC_TASD", "1") != "0":
            try:
                hollow_before = state_dict.is_hollow
                payload = None if hollow_before else list(state_dict.pop_tensors())
            except NotImplementedError:
                payload = None
            if payload is not None:
                devices = {t.get_device() for t in payload if t.is_cuda}
                if len(devices) == 1 and all(t.is_cuda for t in payload):
                    from ...b200.engine import SnapshotEngine

                    snap = SnapshotEngine.get(devices.pop()).snapshot(payload)
                    state_dict.insert_tensors(snap.host_views())
                    return snap
                state_dict.insert_tensors(payload)  # host or mixed payload: leave it to the state dict
        result = state_dict.copy_tensors_to_cpu(non_blocking=True)
        return result if hasattr(result, "descriptor") else None

    @debug_time("BaseCheckpointManager.save", logger)
    def save(self, state_dict: TensorAwareStateDict, iteration: int, is_async: bool = False) -> Optional[AsyncRequest]:
        """Save ``state_dict`` as ``iteration``.  Collective.

        ``is_async=True`` returns an ``AsyncRequest`` to hand to an ``AsyncCallsQueue``; otherwise the save,
        the barrier and the finalization happen inline.  ``state_dict`` is modified: its tensors become host
        copies (or it is left hollow when replication is on)."""
        assert (
            self.latest_iteration < iteration
        ), f"A newer checkpoint is already available: {self.latest_iteration} (saving {iteration})"
        my_id = self._ckpt_id(iteration)
        self._reap_abandoned()
        snaps = []
        if self.repl_strategy:
            replicas, ids = self.repl_strategy.replicate(state_dict, my_id)
            to_save = dict(zip(ids, replicas))
            snaps.extend(getattr(self.repl_strategy, "pop_snapshots", lambda: [])())
            snap = self._snapshot_state_dict(to_save[my_id])
        else:
            to_save = {my_id: state_dict}
            snap = self._snapshot_state_dict(state_dict)
        if snap is not None:
            snaps.append(snap)
        # no engine snapshot at all: the state dict made its own (reference-style) host copies -> device sync
        needs_sync = not snaps or _eager_sync()
        descs = tuple(s.descriptor() for s in snaps)
        self.latest_iteration = -1  # cache invalid from here on

        @debug_time("finalize_fn", logger)
        def finalize_fn():
            executor = ThreadPoolExecutor(max_workers=1)
            validated = self.find_latest()
            self.latest_iteration = -1
            for s in snaps:
                s.release()
            if validated < iteration:
                if is_async:
                    executor.submit(self._cleanup_failed_save, iteration)
                    executor.shutdown(wait=False)
                else:
                    self._cleanup_failed_save(iteration)
                raise CheckpointingException(
                    f"Failure during saving local checkpoint from iteration {iteration}"
                    f" (last valid iteration is {validated})"
                )
            if validated == iteration:
                logging.info(f"Successfully saved local checkpoint from iteration {iteration}")
            else:
                logger.warning(
                    f"WARNING: during saving iteration {iteration} found valid checkpoint from iteration {validated}"
                )
            if is_async:
                executor.submit(self._cleanup, iteration)
                executor.shutdown(wait=False)
            else:
                self._cleanup(iteration)

        if needs_sync and torch.cuda.is_available():
            with debug_time("ckpt_D2H_synchronize", logger):
                if not snaps:
                    torch.cuda.synchronize()
                for s in snaps:
                    s.wait()
        if is_async:
            request = AsyncRequest(self._save_fn, (to_save, descs), [finalize_fn], async_fn_kwargs={})
            self._track(request, snaps)
            return request

        try:
            self._save_fn(to_save, descs)
        except BaseException:
            for s in snaps:  # a failed synchronous save must not keep its pinned slots
                s.release()
            raise
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
        finalize_fn()

    def _track(self, request: AsyncRequest, snaps) -> None:
        """Remember the host slots of an asynchronous save until its ``finalize_fn`` has run -- or can never run any more.

        ``finalize_fn`` releases the slots.  A save whose queue was aborted (``abort_nvrx_checkpoint`` -- the in-process restart
        path, reference ``inprocess/abort.py:201``) or whose request the caller dropped never runs it, and each such save would pin
        one or two snapshot-sized shm slots for good.  The request is the only owner of its ``finalize_fn``: when the queue and
        the caller have let go of the request, the function object dies, and that is the moment the slots go back (releasing is
        idempotent, so the normal path is unaffected)."""
        import weakref

        def give_back(_ref, snaps=tuple(snaps)):
            for s in snaps:
                s.release()

        self._outstanding.append((weakref.ref(request.finalize_fns[0], give_back), snaps))

    def _reap_abandoned(self):
        """Forget the saves that are done with their slots (finalized, or abandoned and given back by :meth:`_track`)."""
        self._outstanding = [(ref, snaps) for ref, snaps in self._outstanding if not all(s.released for s in snaps)]

    def release_unfinalized(self):
        """Release the host slots of every save of this manager that has not been finalized (call after aborting the queue,
        e.g. from an in-process restart handler).  The saves themselves are lost, like in the reference."""
        for _, snaps in self._outstanding:
            for s in snaps:
                s.release()
        self._outstanding = []

    def __del__(self):
        try:
            self.release_unfinalized()
        except Exception:  # noqa: BLE001
            pass
