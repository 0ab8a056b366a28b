"""``FileSystemWriterAsync``: torch.distributed.checkpoint (DCP) storage writer split into plan / stage / write.

Entry points mirror reference ``checkpointing/async_ckpt/filesystem_async.py`` (``FileSystemWriterAsync`` ``:140``,
``prepare_write_data`` ``:214``, ``get_save_function_and_args`` ``:506``, ``retrieve_write_results`` ``:1174``,
``get_write_results_queue`` ``:120``) so the call sequence of ``examples/checkpointing/async_writer.py`` is unchanged:

    writer = FileSystemWriterAsync(dir, thread_count=2)
    ret = save_state_dict_async_plan(state_dict, writer, None, 0, planner=planner)
    save_fn, preload_fn, save_args = writer.get_save_function_and_args()
    queue.schedule_async_request(AsyncRequest(save_fn, save_args, [finalize], preload_fn=preload_fn))

What differs is the staging.  The reference resolves every write item and copies each CUDA tensor to
the host on its own (``preload_tensors`` ``:564-691``, through CUDA-IPC handles or per-tensor shared-memory tensors).  Here
``prepare_write_data`` hands ALL CUDA tensors of the plan to the snapshot engine: one pack kernel, one side-stream drain into
one pinned shared-memory slot.  The write function (run by the async worker) maps the slot, waits for the drain on its
progress word, and then lets PyTorch's own ``FileSystemWriter.write_data`` produce the files from the host views -- the
on-disk format is PyTorch's, byte for byte (``tests/test_dcp_async_cpu.py`` compares against a synchronous ``dcp.save``).

``separation_hint`` works as in the reference (``:1350``): items whose FQN starts with the hint go to files of their own,
named ``<hint>__<rank>_<n>.distcp``, and ``thread_count`` must then be at least 2.  Quantized CUDA tensors that offer
``dequantize()`` are dequantized before staging (``:253``).

Not carried over (accepted for signature compatibility, documented no-ops): ``use_msc`` (multistorageclient, raises),
``is_multiproc_io`` (file IO is multi-threaded through ``thread_count``), ``use_cached_data_structure`` and
``use_cpu_shm_for_gpu_tensors`` (the engine's plan cache and shm slot make both moot).
"""

from __future__ import annotations

import logging
import os
import queue as queue_mod
from functools import partial
from time import time
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch
from torch import multiprocessing as mp
from torch.distributed.checkpoint import FileSystemWriter
from torch.distributed.checkpoint.planner import SavePlan, SavePlanner, WriteItem, WriteItemType
from torch.distributed.checkpoint.storage import WriteResult
from torch.distributed.checkpoint.utils import _is_wrapped_exception, _wrap_exception

from ..b200.persist import drain_aware

logger = logging.getLogger(__name__)

WRAPPED_EXCEPTION = Tuple[BaseException, object]

_results_queue = None
_results_manager = None
_live_saves: set = set()  # ids of saves of this process whose results have not been retrieved (or given up) yet


def get_write_results_queue(mp_mode: str = "spawn"):
    """Process-wide queue the writer process reports ``(rank, results-or-exception)`` on (a manager queue, so it can
    travel inside an ``AsyncRequest`` to the persistent worker)."""
    global _results_queue, _results_manager
    if _results_queue is None:
        _results_manager = mp.get_context(mp_mode).Manager()
        _results_queue = _results_manager.Queue()
    return _results_queue


class _HostPlanner:
    """The slice of the SavePlanner interface ``FileSystemWriter.write_data`` uses, answering from staged host data."""

    def __init__(self, staged: Dict):
        self._staged = staged

    def resolve_data(self, write_item: WriteItem):
        return self._staged[write_item.index]


def _passthrough(payload):
    return payload


def _use_open_file(writer: FileSystemWriter, open_file: Callable) -> None:
    """Make ``writer`` open its data files through the user's ``open_file(path, mode)``."""
    from contextlib import contextmanager

    @contextmanager
    def create_stream(path, mode):
        with open_file(os.fspath(path), mode) as stream:
            yield stream

    writer.fs.create_stream = create_stream


def _write_separated(writer: FileSystemWriter, plan: SavePlan, planner, hint: str):
    """``FileSystemWriter.write_data`` with the items split by ``hint``: half as many size-balanced buckets as threads, each
    bucket divided into "FQN starts with the hint" and the rest, one file per non-empty part; files of the hinted part carry
    the hint in front of the usual ``__<rank>_`` prefix."""
    from torch.distributed.checkpoint.filesystem import DEFAULT_SUFFIX, _split_by_size_and_type

    buckets = _split_by_size_and_type(max(1, writer.thread_count // 2), plan.items)
    files: "queue_mod.Queue" = queue_mod.Queue()
    count = 0
    for group in ("", hint):
        for bucket in buckets:
            part = [it for it in bucket if it.index.fqn.startswith(hint) == bool(group)]
            if part:
                name = f"{group}{plan.storage_data.prefix}{count}{DEFAULT_SUFFIX}"
                count += 1
                files.put((writer.fs.concat_path(writer.path, name), name, part))
    return writer._write_data(planner, files)


class FileSystemWriterAsync(FileSystemWriter):
    """Async-capable DCP filesystem writer.  One instance per checkpoint save (state is kept between the stages).

    Flow: ``prepare_write_data`` (trainer: resolve + snapshot) -> ``get_save_function_and_args`` -> the returned function
    runs in a writer process -> ``retrieve_write_results`` -> ``finish`` (coordinator writes ``.metadata``)."""

    # Kept for API compatibility (reference ``filesystem_async.py:161,167``; its tests clear them between cases): the reference
    # caches "which structures the worker already holds" and its shm staging tensors per writer class; here staging belongs to
    # the snapshot engine (plans per structure, pooled slots), so both stay empty.
    _cached_identifiers: set = set()
    _shm_tensor_cache: Dict = {}

    def __init__(
        self,
        path: Union[str, os.PathLike],
        *args,
        separation_hint: Optional[str] = None,
        use_msc: bool = False,
        is_multiproc_io: bool = False,
        use_cached_data_structure: bool = False,
        use_cpu_shm_for_gpu_tensors: bool = False,
        **kwargs,
    ):
        if use_msc:
            raise NotImplementedError("use_msc (multistorageclient) is not supported by the B200 writer")
        self.checkpoint_dir = path
        self.use_msc = use_msc
        # ``open_file(path, mode)`` (reference ``:214``; its tests inject a failing one to see errors reported): used by the
        # writer process for this rank's data files instead of PyTorch's own open
        self.open_file = kwargs.pop("open_file", None)
        if self.open_file is open:
            self.open_file = None
        super().__init__(path, *args, **kwargs)
        if not self.single_file_per_rank:
            raise NotImplementedError("single_file_per_rank flag not supported for FileSystemWriterAsync")
        self._ctor = (os.fspath(path), args, dict(kwargs))
        self.can_run_decentralized_global_plan: bool = True
        self.separation_hint = separation_hint
        self.is_multi_proc_io = is_multiproc_io
        self.use_cached_data_structure = use_cached_data_structure
        self.use_cpu_shm_for_gpu_tensors = use_cpu_shm_for_gpu_tensors
        self.has_data_to_write: bool = False
        self.results_queue = None
        self._payload: Optional[dict] = None
        self._snapshot = None
        self._save_id: Optional[str] = None

    # ---- stage 1 (trainer) ------------------------------------------------------------------------
    def prepare_write_data(self, plan: SavePlan, planner: SavePlanner) -> None:
        """Resolve every write item; CUDA tensors are snapshotted together (pack kernel + drain into one shm slot), host
        tensors and byte blobs are kept as they are.  Returns after the GPU work is *enqueued*."""
        t0 = time()
        if self.separation_hint:
            assert self.thread_count > 1, "thread_count must be at least 2 if separation_hint is provided"
        staged_host: Dict = {}
        cuda_items: List[WriteItem] = []
        cuda_tensors: List[torch.Tensor] = []
        for item in plan.items:
            data = planner.resolve_data(item)
            if item.type == WriteItemType.BYTE_IO:
                staged_host[item.index] = data.getvalue() if hasattr(data, "getvalue") else bytes(data)
            else:
                ten = data.detach()
                if ten.is_cuda:
                    if "dequantize" in type(ten).__dict__:  # quantized wrapper tensors are stored dequantized
                        ten = ten.dequantize()
                    cuda_items.append(item)
                    cuda_tensors.append(ten)
                else:
                    # host tensors are written as they are *now* (the reference clones them for the same reason)
                    staged_host[item.index] = ten.clone()
        desc = None
        if cuda_tensors:
            from ..b200.engine import SnapshotEngine

            devices = {t.device.index for t in cuda_tensors}
            assert len(devices) == 1, f"write items on several CUDA devices: {sorted(devices)}"
            self._snapshot = SnapshotEngine.get(devices.pop()).snapshot(cuda_tensors)
            desc = self._snapshot.descriptor()
        self._payload = {
            "plan": plan, "host": staged_host, "snapshot": desc, "cuda_indices": [it.index for it in cuda_items],
        }
        self.has_data_to_write = bool(plan.items)
        logger.debug(f"prepare_write_data: {len(plan.items)} items ({len(cuda_tensors)} on the GPU) in {time() - t0:.4f}s")

    def get_save_function_and_args(self) -> Tuple[Optional[Callable], Optional[Callable], List]:
        """``(save_fn, preload_fn, args)`` for an ``AsyncRequest``; ``args[1]`` is the payload slot ``preload_fn`` fills.

        ``preload_fn`` returns immediately (the staging was enqueued by ``prepare_write_data`` and the write function waits
        for it in the writer process), so the trainer is not held for the D2H as in the reference (``core.py:547``)."""
        if not self.has_data_to_write:
            return None, None, []
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.results_queue = get_write_results_queue()
        # the results queue is shared by every save of this process; a save that was aborted after its writer had reported
        # leaves an entry behind that nobody asks for -- entries carry the id of their save so that it is never mistaken for
        # the result of a later one
        import uuid

        self._save_id = uuid.uuid4().hex
        _live_saves.add(self._save_id)
        options = {"open_file": self.open_file, "multiproc": bool(self.is_multi_proc_io)}
        save_fn = drain_aware(
            partial(self.write_preloaded_data, self._ctor, int(self.thread_count), self.separation_hint, (self._save_id, options))
        )
        return save_fn, partial(_passthrough, self._payload), [rank, None, self.results_queue]

    # ---- stage 2 (writer process) -------------------------------------------------------------------
    @staticmethod
    def write_preloaded_data(ctor, thread_count: int, separation_hint, save_id, rank: int, payload: dict, results_queue) -> None:
        """Write the files of this rank with PyTorch's ``FileSystemWriter`` from staged host data; the outcome (list of
        ``WriteResult`` or a wrapped exception) is reported on ``results_queue``.  Errors are reported, not raised; only a
        ``SystemExit`` / ``KeyboardInterrupt`` (the worker is being aborted) passes through, without a report."""
        outcome = None
        held = []
        options = {}
        if isinstance(save_id, tuple):
            save_id, options = save_id
        try:
            if options.get("multiproc"):
                import multiprocessing

                if multiprocessing.current_process().daemon:
                    # the reference forks one process per file bucket in this mode (``:805-815``), which a daemonic worker
                    # may not do; the wording is the reference's.  (Here the files are written by threads of PyTorch's
                    # FileSystemWriter and, for snapshot payloads, by the slot's writer pool -- no extra processes.)
                    raise RuntimeError("Invalid Setup! User cannot establish a daemon Async worker and then use Multi-Proc File IO.")
            staged = dict(payload["host"])
            for key, blob in list(staged.items()):
                if isinstance(blob, (bytes, bytearray)):
                    import io

                    staged[key] = io.BytesIO(blob)
            if payload["snapshot"] is not None:
                from ..b200.engine import host_views
                from ..b200.persist import DRAIN_TIMEOUT_MS, open_slot

                hb = open_slot(payload["snapshot"]["shm_name"], cache=os.environ.get("NVRX_B200_CACHE_SLOTS") == "1")
                held.append(hb)
                hb.wait(payload["snapshot"]["progress_target"], DRAIN_TIMEOUT_MS)
                for index, view in zip(payload["cuda_indices"], host_views(payload["snapshot"]["layout"], hb)):
                    staged[index] = view
            path, args, kwargs = ctor
            writer = FileSystemWriter(path, *args, **kwargs)
            writer.thread_count = thread_count
            # everything is on the host already: keep PyTorch off its CUDA copy-ahead loader (it would create a CUDA
            # context in the writer process, or fail in a forked one)
            writer.per_thread_copy_ahead = 0
            if options.get("open_file") is not None:
                _use_open_file(writer, options["open_file"])
            if separation_hint:
                outcome = _write_separated(writer, payload["plan"], _HostPlanner(staged), separation_hint).wait()
            else:
                outcome = writer.write_data(payload["plan"], _HostPlanner(staged)).wait()
            if outcome is None:
                outcome = []
        except Exception as exc:  # noqa: BLE001 - reported to the trainer, raised there on the coordinator
            logger.error(f"rank {rank}: checkpoint write failed: {exc}", exc_info=True)
            outcome = _wrap_exception(exc)
        finally:
            for hb in held:
                hb.close(unlink=False)
        results_queue.put((rank, save_id, outcome))

    # ---- stage 3 (trainer) ------------------------------------------------------------------------
    def write_data(self, plan: SavePlan, planner: SavePlanner):
        raise NotImplementedError("FileSystemWriterAsync splits write_data into prepare_write_data + write_preloaded_data")

    def retrieve_write_results(self) -> Union[List[WriteResult], WRAPPED_EXCEPTION]:
        """Results of this rank's write (or the wrapped exception it died with); releases the snapshot slot."""
        try:
            if not self.has_data_to_write or self.results_queue is None:
                return []
            rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
            stash = []
            try:
                while True:
                    got_rank, save_id, outcome = self.results_queue.get(timeout=600)
                    if save_id == self._save_id:
                        break
                    if save_id in _live_saves:
                        stash.append((got_rank, save_id, outcome))  # another save of this process still in flight: put it back
                    else:
                        logger.debug(f"rank {rank}: dropping the result of an abandoned save ({save_id})")
            except queue_mod.Empty:
                return _wrap_exception(RuntimeError(f"rank {rank}: no results from the checkpoint writer"))
            finally:
                for item in stash:
                    self.results_queue.put(item)
            if _is_wrapped_exception(outcome):
                # same wording as the reference (``filesystem_async.py:1199``) so callers matching on it keep working
                try:
                    raise RuntimeError(f"Worker failure: {outcome[0]}") from outcome[0]
                except RuntimeError as exc:
                    return _wrap_exception(exc)
            if self.has_data_to_write and len(outcome) == 0:
                return _wrap_exception(RuntimeError(f"rank {rank}: the writer reported no results for a non-empty plan"))
            return outcome
        finally:
            _live_saves.discard(self._save_id)
            if self._snapshot is not None:
                self._snapshot.release()
                self._snapshot = None
            self._payload = None

    def __del__(self):
        # a save that was aborted (in-process restart) never reaches retrieve_write_results: give its host slot back and mark
        # whatever its writer may still report as nobody's
        try:
            _live_saves.discard(getattr(self, "_save_id", None))
        except Exception:  # noqa: BLE001 - module globals are gone at interpreter shutdown
            pass
        snap, self._snapshot = getattr(self, "_snapshot", None), None
        if snap is not None:
            try:
                snap.release()
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass

    def prepare_decentralized_global_plan(self, local_plan: SavePlan) -> SavePlan:
        """Storage prefix for planning without a reduce_scatter: files of rank r start with ``__r_``."""
        import dataclasses

        from torch.distributed.checkpoint.filesystem import _StoragePrefix

        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        return dataclasses.replace(local_plan, storage_data=_StoragePrefix(f"__{rank}_"))

    def finish(self, metadata, results: List[List[WriteResult]]) -> None:
        """Coordinator: write ``.metadata``.  The saver hangs every rank's local plan on the metadata object
        (``all_local_plans``) so that a job resuming from this checkpoint can skip the plan exchange; recent PyTorch copies the
        dataclass inside ``finish`` (``dataclasses.replace``) and loses such extra attributes, so they are put back into the
        written file (small, atomic rewrite).  Readers that do not know the attribute ignore it."""
        plans = getattr(metadata, "all_local_plans", None)
        super().finish(metadata, results)
        if plans is None:
            return
        import pickle

        path = os.path.join(os.fspath(self.checkpoint_dir), ".metadata")
        try:
            with open(path, "rb") as fh:
                written = pickle.load(fh)  # nosec B301 - the file this process has just written
            if getattr(written, "all_local_plans", None) is None:
                written.all_local_plans = plans
                tmp = path + f".plans{os.getpid()}"
                with open(tmp, "wb") as fh:
                    pickle.dump(written, fh)
                os.replace(tmp, path)
        except OSError as exc:  # the checkpoint is complete without it; only the resume shortcut is lost
            logger.warning(f"could not store the local plans in {path}: {exc}")

    @staticmethod
    def preload_tensors(payload, non_blocking: bool = True):
        """Kept for API compatibility (reference ``:563``): there the D2H of the write items happens here; in this
        implementation it was enqueued by ``prepare_write_data`` already, so the payload passes through."""
        return payload

    @property
    def checkpoint_id(self) -> Union[str, os.PathLike]:
        return self.checkpoint_dir

    @classmethod
    def cleanup_tensor_caches(cls) -> None:
        """Kept for API compatibility: staging lives in the snapshot engine, there is no per-writer tensor cache."""

    @classmethod
    def register_shm_drain_callback(cls, fn: Optional[Callable[[], None]]) -> None:
        """Kept for API compatibility: slot reuse is fenced by the engine (a slot is busy until its save is finalized)."""
