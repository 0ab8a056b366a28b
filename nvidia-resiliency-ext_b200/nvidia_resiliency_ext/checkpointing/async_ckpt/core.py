"""Async-call scheduling for checkpoint saves (API mirror of reference ``checkpointing/async_ckpt/core.py``).

Same public surface -- ``AsyncRequest`` (``:120``), ``AsyncCaller`` (``:232``), ``TemporalAsyncCaller`` (``:308``, fork),
``PersistentAsyncCaller`` (``:428``, spawn + queues), ``AsyncCallsQueue`` (``:841``), ``abort_nvrx_checkpoint``
(``:1041``) -- so NeMo / Megatron callers are unchanged.  Rebuilt pieces, all on the snapshot hand-off:

* **no device-wide sync for engine snapshots.**  The reference callers run ``torch.cuda.synchronize()`` before
  handing the request to the writer (``:345``) because the writer reads pinned tensors whose D2H must have
  finished.  Requests whose ``async_fn`` is *drain aware* (attribute ``nvrx_drain_aware``; see
  ``checkpointing/b200/persist.py``) carry a shared-memory snapshot descriptor instead: the writer process
  follows the drain through the buffer's progress word, so the trainer returns right after the pack kernel
  is enqueued.  Any other request keeps the reference behaviour (sync before fork).
* **no tensor payload through the mp queue.**  A drain-aware request pickles to a few KB (shm name + layout);
  the reference ships every CPU tensor through ``torch.multiprocessing`` reductions, i.e. one more full
  host copy of the snapshot (survey appendix A.12).
* **blocking finalize wakes on completion** (``Queue.get(timeout)``) instead of the reference's 100 ms sleep
  poll (``:589``), and notices a dead worker instead of spinning forever (``:577-589``).
"""

from __future__ import annotations

import gc
import logging
import os
import shutil
import signal
import subprocess  # nosec B404
import weakref
from abc import ABC, abstractmethod
from collections import deque
from queue import Empty
from time import time
from typing import Callable, ClassVar, Dict, List, NamedTuple, Optional, Tuple

import torch
from torch import multiprocessing as mp

from ..utils import _disable_gc, debug_time

logger = logging.getLogger(__name__)

_POLL_SLICE_S = 0.02  # upper bound on how long a blocking wait sleeps between liveness checks


def _set_process_qos(cpu_priority: int, io_priority: Optional[int]) -> None:
    """De-prioritise the calling (writer) process: ``nice`` to ``cpu_priority`` (0-19, only ever raised)
    and optionally ``ionice -c io_priority`` (0-3; 3 = idle is the useful value).  Failures are logged,
    never raised."""
    pid = os.getpid()
    if cpu_priority is not None and 0 <= cpu_priority <= 19:
        try:
            now = os.nice(0)
            if cpu_priority > now:
                logger.debug("PID %s: nice %s -> %s", pid, now, os.nice(cpu_priority - now))
            else:
                logger.warning(
                    "PID %s: nice already %s (>= requested %s); lowering needs superuser", pid, now, cpu_priority
                )
        except (OSError, PermissionError) as exc:
            logger.warning(f"PID {pid}: could not change CPU priority: {exc}")
    if io_priority is None:
        return
    if io_priority not in (0, 1, 2, 3):
        logger.warning(f"PID {pid}: io_priority {io_priority!r} is not an ionice class (0-3); ignored")
        return
    if io_priority <= 2:
        logger.warning(
            f"PID {pid}: io_priority={io_priority} does not de-prioritise I/O (1 = realtime raises it, "
            f"2 = best-effort is the default); use 3 (idle) for checkpoint writers"
        )
    try:
        tool = shutil.which("ionice")
        if tool is None:
            raise FileNotFoundError("ionice not on PATH")
        subprocess.run([tool, "-c", str(io_priority), "-p", str(pid)], check=True, capture_output=True)  # nosec B603
        logger.debug(f"PID {pid}: ionice class {io_priority}")
    except (subprocess.CalledProcessError, FileNotFoundError, PermissionError) as exc:
        logger.warning(f"PID {pid}: could not change I/O priority: {exc}")


def _collective_device() -> torch.device:
    """Device for the tiny agreement all-reduces: the GPU under NCCL, the CPU under gloo."""
    if torch.cuda.is_available() and torch.distributed.get_backend() != "gloo":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _is_drain_aware(fn: Optional[Callable]) -> bool:
    return bool(getattr(fn, "nvrx_drain_aware", False))


class AsyncRequest(NamedTuple):
    """One asynchronous save.

    Fields (identical to the reference NamedTuple, positional order matters to callers):
        async_fn:        function run by the writer process; ``None`` = nothing to run
        async_fn_args:   its positional arguments; **index 1 is the payload slot** -- when ``preload_fn`` is set
                         its return value replaces ``async_fn_args[1]``
        finalize_fns:    run on the trainer, in order, after ``async_fn`` finished on all ranks
        async_fn_kwargs: keyword arguments of ``async_fn``
        preload_fn:      optional staging function (device -> host) executed before ``async_fn``
        is_frozen:       frozen requests reject new finalize functions
        call_idx:        sequence number assigned by the queue
    """

    async_fn: Optional[Callable]
    async_fn_args: Tuple
    finalize_fns: List[Callable]
    async_fn_kwargs: Optional[Dict] = None
    preload_fn: Callable = None
    is_frozen: bool = False
    call_idx: int = 0

    def add_finalize_fn(self, fn: Callable) -> None:
        """Append ``fn`` after the finalize functions already registered."""
        if self.is_frozen:
            raise RuntimeError("Cannot add finalization functions to a frozen AsyncRequest")
        self.finalize_fns.append(fn)

    def _resolved_args(self) -> list:
        args = list(self.async_fn_args)
        if self.preload_fn:
            args[1] = self.preload_fn()
        return args

    def execute_sync(self) -> None:
        """Run the whole request inline: preload, save, barrier, finalize (what the async path amounts to)."""
        args = self._resolved_args()
        if self.async_fn is not None:
            self.async_fn(*args, **dict(self.async_fn_kwargs or {}))
        torch.distributed.barrier()
        self.execute_finalize_fns(validate_matching_call_idx=False)

    def freeze(self) -> "AsyncRequest":
        """Copy of the request that no longer accepts finalize functions."""
        return self._replace(is_frozen=True)

    def execute_finalize_fns(self, validate_matching_call_idx: bool = True) -> int:
        """Run the finalize functions; optionally check (one int32 MAX all-reduce) that every rank is
        finalizing the same ``call_idx``.  Returns ``call_idx``."""
        with debug_time("finalize", logger):
            for fn in self.finalize_fns:
                fn()
            if validate_matching_call_idx:
                probe = torch.tensor([self.call_idx], dtype=torch.int, device=_collective_device())
                torch.distributed.all_reduce(probe, op=torch.distributed.ReduceOp.MAX)
                assert probe.item() == self.call_idx, (
                    "Unmatched async calls. That probably means not all ranks are participating in async finalization"
                )
        return self.call_idx


class ObjectTracker(type):
    """Metaclass keeping weak references to every instance (``get_instances``), used by the abort path."""

    def __init__(cls, name, bases, attrs):
        super().__init__(name, bases, attrs)
        cls._instances = weakref.WeakSet()

    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        cls._instances.add(obj)
        return obj

    def get_instances(cls):
        return list(cls._instances)


class AsyncCaller(ABC):
    """Owns the writer process of one (or, for the persistent flavour, all) async request(s) and answers
    "is the current call done on every rank?"."""

    def __init__(self):
        self.process: Optional[mp.Process] = None
        self.start_time: Optional[float] = None
        self.rank: int = None  # remembered for logging after torch.distributed is gone

    @abstractmethod
    def schedule_async_call(self, async_req: AsyncRequest) -> None:
        """Hand ``async_req`` to a writer process.  Collective: call on all ranks."""
        raise NotImplementedError

    @abstractmethod
    def is_current_async_call_done(self, blocking: bool, no_dist: bool) -> bool:
        """True when the current call finished (on all ranks unless ``no_dist``); ``blocking`` waits for it."""
        raise NotImplementedError

    def sync_all_async_calls(self, is_alive: int) -> bool:
        """Sum ``is_alive`` over ranks; done iff nobody is still writing."""
        flag = torch.tensor([is_alive], dtype=torch.int, device=_collective_device())
        torch.distributed.all_reduce(flag)
        return flag[0] == 0

    @abstractmethod
    def close(self, abort=False):
        """Terminate the caller; ``abort=True`` kills instead of waiting."""
        raise NotImplementedError

    @abstractmethod
    def __del__(self):
        raise NotImplementedError

    def _remember_rank(self) -> None:
        if self.rank is None:
            self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0


class TemporalAsyncCaller(AsyncCaller):
    """One forked child per request (the flavour ``LocalCheckpointManager`` uses; reference ``:308-425``).

    The child inherits the request by copy-on-write and must not touch CUDA.  For a drain-aware
    ``async_fn`` the fork happens immediately (the child waits for the drain on the shared progress word);
    otherwise the device is synchronised first, as in the reference (``:345``)."""

    def __init__(self):
        super().__init__()
        self.preloaded_holder = None

    @_disable_gc()
    def schedule_async_call(self, async_req: AsyncRequest) -> None:
        if async_req.async_fn is None:
            return
        args = list(async_req.async_fn_args)
        if async_req.preload_fn:
            args[1] = async_req.preload_fn()
            self.preloaded_holder = args[1]  # keep host staging alive until the child is joined
        self._remember_rank()

        if not _is_drain_aware(async_req.async_fn) and torch.cuda.is_available():
            t0 = time()
            torch.cuda.synchronize()
            logger.debug(f"rank: {self.rank}, takes {time() - t0} to finish D2H ")

        self.start_time = time()
        self.process = mp.get_context("fork").Process(
            target=async_req.async_fn, args=args, kwargs=dict(async_req.async_fn_kwargs or {})
        )
        self.process.start()
        logger.debug(f"rank: {self.rank}, takes {time() - self.start_time} to schedule async ckpt ")

    def is_current_async_call_done(self, blocking: bool = False, no_dist: bool = False) -> bool:
        alive = int(self.process.is_alive()) if self.process is not None else 0
        done = (not alive) if no_dist else self.sync_all_async_calls(alive)
        if done or blocking:
            self.close()  # joins (waits) when blocking
            done = True
        return done

    def close(self, abort=False):
        if not self.process:
            return
        if abort:
            logger.warning(f"Temporal worker aborted in rank {self.rank}")
            self.process.kill()
            self.process.join()
        else:
            self.process.join()
            if self.process.exitcode not in (0, None):
                logger.error(f"rank {self.rank}: async writer exited with code {self.process.exitcode}")
        logger.debug(f"TemporalAsyncCaller: writer joined {time() - self.start_time:.2f}s after the fork")
        self.process = None
        self.start_time = None
        self.preloaded_holder = None

    def __del__(self):
        pass

    def _debug_is_async_process_running(self):
        """Test hook: is the forked writer alive?"""
        return self.process is not None and self.process.is_alive()


class PersistentAsyncCaller(AsyncCaller):
    """One long-lived spawned worker fed through queues (reference ``:428-823``).

    ``queue``      trainer -> worker: ``AsyncRequest`` objects and the ``'DONE'`` sentinel
    ``preload_q``  hand-shake around ``preload_fn`` (the trainer blocks until staging finished)
    ``comp_q``     worker -> trainer: ``call_idx`` of every finished request
    """

    _worker_data_cache: Dict = {}  # worker-side cache (e.g. IPC handles of the DCP writer); cleared on exit
    _worker_restart_callbacks: ClassVar[List[Callable]] = []

    @classmethod
    def register_worker_restart_callback(cls, fn: Callable) -> None:
        """``fn()`` runs in the trainer every time a fresh worker is spawned."""
        cls._worker_restart_callbacks.append(fn)

    def __init__(
        self,
        is_daemon: bool = True,
        cpu_priority: int = 10,
        io_priority: Optional[int] = None,
        sigterm_timeout: float = 30.0,
        cpu_shm_mode: bool = False,
    ):
        self.process: mp.Process = None
        self.start_time: Optional[float] = None
        self.sigterm_timeout = sigterm_timeout
        ctx = mp.get_context("spawn")
        self.queue: mp.JoinableQueue = ctx.JoinableQueue()
        self.preload_q: mp.JoinableQueue = ctx.JoinableQueue()
        self.comp_q: mp.Queue = ctx.Queue()
        self.cur_item: int = None
        self.cur_idx: int = -1
        self.rank: int = None
        self.background_worker_is_daemon = is_daemon
        self.cpu_priority = cpu_priority
        self.io_priority = io_priority
        self.cpu_shm_mode = cpu_shm_mode

    def _start_worker(self, rank: int) -> None:
        ctx = mp.get_context("spawn")
        logger.info(f"PersistentAsyncCaller: {rank}, Starting Async Caller")
        target = (
            PersistentAsyncCaller.async_loop_for_daemon_worker
            if self.background_worker_is_daemon
            else PersistentAsyncCaller.async_loop
        )
        self.process = ctx.Process(
            target=target,
            args=(
                rank, self.queue, self.preload_q, self.comp_q, logger.getEffectiveLevel(),
                self.cpu_priority, self.io_priority, self.cpu_shm_mode,
            ),
            daemon=self.background_worker_is_daemon,
        )
        self.process.start()
        logger.debug(f"PersistentAsyncCaller: {rank}, Started Async Caller {self.process}")
        for cb in PersistentAsyncCaller._worker_restart_callbacks:
            cb()

    def schedule_async_call(self, async_req: AsyncRequest) -> None:
        """Queue the (picklable) request; blocks only while a ``preload_fn`` runs in the worker."""
        if async_req.async_fn is None:
            return
        self._remember_rank()
        self.start_time = time()
        if self.process is None:
            self._start_worker(self.rank)
        if async_req.preload_fn:
            self.preload_q.put(async_req.call_idx)
        self.queue.put(async_req)
        logger.debug(f"rank: {self.rank}, put {async_req.call_idx}")
        if async_req.preload_fn:
            t0 = time()
            self.preload_q.join()
            logger.debug(f"rank: {self.rank}, takes {time() - t0} to finish D2H ")
        logger.debug(f"rank: {self.rank}, takes {time() - self.start_time} to schedule async ckpt ")

    def _wait_completion(self, blocking: bool) -> bool:
        """Fetch the next completed ``call_idx`` into ``cur_item``.  Returns True while still running."""
        while self.cur_item is None:
            try:
                self.cur_item = self.comp_q.get(timeout=_POLL_SLICE_S) if blocking else self.comp_q.get_nowait()
            except Empty:
                if not blocking:
                    return True
                if self.process is not None and not self.process.is_alive():
                    # one last look: the worker may have finished the item right before exiting
                    try:
                        self.cur_item = self.comp_q.get(timeout=_POLL_SLICE_S)
                    except Empty:
                        raise RuntimeError(
                            f"rank {self.rank}: persistent checkpoint worker died "
                            f"(exit code {self.process.exitcode}) with a save in flight"
                        ) from None
        return False

    def is_current_async_call_done(self, blocking: bool = False, no_dist: bool = False) -> bool:
        running = self._wait_completion(blocking) if self.process else False
        if self.cur_item is not None:
            logger.debug(f"rank: {self.rank}, item: {self.cur_item} is completed, {running}")
        done = (not running) if no_dist else self.sync_all_async_calls(int(running))
        if done:
            logger.debug(f"rank: {self.rank}, item: {self.cur_item} is completed globally, {done}")
            self.cur_item = None
        return done

    def close(self, abort=False):
        """Drain and stop the worker (``'DONE'`` sentinel), or abort it: SIGTERM, then SIGKILL after
        ``sigterm_timeout`` seconds with the queues closed first so no IPC state dangles."""
        logger.info(f"PersistentAsyncCaller: {self.rank}, Destroying Async Caller")
        if not self.process:
            return
        if abort:
            logger.error(f"Persistent worker aborted in rank {self.rank}")
            self.process.terminate()
            self.process.join(timeout=self.sigterm_timeout)
            if self.process.is_alive():
                logger.warning(
                    f"Persistent worker (rank {self.rank}) ignored SIGTERM for {self.sigterm_timeout}s; sending SIGKILL"
                )
                for q in (self.queue, self.preload_q):
                    try:
                        q.cancel_join_thread()
                        q.close()
                    except Exception:  # noqa: BLE001
                        pass
                gc.collect()
                self.process.kill()
                self.process.join()
        elif self.process.is_alive():
            self.queue.put("DONE")
            self.queue.join()
            self.process.join()
        else:
            self.process.join()
        self.process = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _debug_is_async_process_running(self):
        """Test hook: is the persistent worker alive?"""
        return self.process is not None and self.process.is_alive()

    @classmethod
    def cleanup_worker_data_cache(cls):
        """Drop worker-side cached structures (they may hold CUDA IPC handles) before the worker exits."""
        if cls._worker_data_cache:
            logger.info(f"Cleaning up worker data cache with {len(cls._worker_data_cache)} entries")
            cls._worker_data_cache.clear()
            gc.collect()
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                torch.cuda.empty_cache()

    @staticmethod
    def async_process_target(
        rank: int,
        queue: mp.JoinableQueue,
        preload_q: mp.JoinableQueue,
        comp_q: mp.Queue,
        log_level: int = logging.INFO,
        cpu_priority: int = 10,
        io_priority: Optional[int] = None,
        cpu_shm_mode: bool = False,
    ):
        """Body of the persistent worker: loop over requests until ``'DONE'``.

        Per request: run ``preload_fn`` (if any) and release the trainer through ``preload_q``; run
        ``async_fn``; report ``call_idx`` on ``comp_q``."""
        logging.getLogger("nvidia_resiliency_ext").setLevel(log_level)
        wlog = logging.getLogger(__name__)
        wlog.log(logging.INFO if rank == 0 else logging.DEBUG, f"PersistentAsyncCaller: persistent ckpt worker for {rank} has started")
        if not cpu_shm_mode and torch.cuda.is_available() and torch.cuda.device_count() > 0:
            # requests of other producers (e.g. the DCP writer) may carry CUDA-IPC tensors: own the right
            # device and force context creation before the first handle arrives
            dev = rank % torch.cuda.device_count()
            torch.cuda.set_device(dev)
            torch.empty(1, device=f"cuda:{dev}")
        _set_process_qos(cpu_priority=cpu_priority, io_priority=io_priority)
        os.environ["NVRX_B200_CACHE_SLOTS"] = "1"  # long-lived writer: keep snapshot slots mapped between saves

        def _on_sigterm(signum, frame):
            raise SystemExit(128 + signum)

        signal.signal(signal.SIGTERM, _on_sigterm)
        try:
            while True:
                item = queue.get()
                if isinstance(item, str) and item == "DONE":
                    queue.task_done()
                    break
                if isinstance(item, AsyncRequest):
                    args = list(item.async_fn_args)
                    if item.preload_fn:
                        staged_idx = preload_q.get()
                        args[1] = item.preload_fn()
                        wlog.debug(f"{rank} has completed D2H of {staged_idx}")
                        preload_q.task_done()
                    if item.async_fn is not None:
                        item.async_fn(*args, **dict(item.async_fn_kwargs or {}))
                    wlog.debug(f"{rank} has completed saving {item.call_idx}")
                    comp_q.put(item.call_idx)
                    queue.task_done()
                    del args
                del item
                gc.collect()
        except RuntimeError as exc:
            if "pidfd_getfd" in str(exc) and "Operation not permitted" in str(exc):
                raise RuntimeError(
                    "The checkpoint worker could not receive a CUDA IPC handle from the trainer (pidfd_getfd: "
                    "Operation not permitted); allow cross-process fd passing: sudo sysctl kernel.yama.ptrace_scope=0"
                ) from exc
            raise
        finally:
            PersistentAsyncCaller.cleanup_worker_data_cache()
        wlog.log(logging.INFO if rank == 0 else logging.DEBUG, f"PersistentAsyncCaller: persistent ckpt worker for {rank} has terminated")

    @staticmethod
    @_disable_gc()
    def async_loop(rank, queue, preload_q, comp_q, log_level=logging.INFO, cpu_priority=10, io_priority=None, cpu_shm_mode=False):
        """Entry point of a non-daemon worker (may spawn children, e.g. parallel file writers)."""
        PersistentAsyncCaller.async_process_target(
            rank, queue, preload_q, comp_q, log_level, cpu_priority, io_priority, cpu_shm_mode
        )

    @staticmethod
    def async_loop_for_daemon_worker(
        rank, queue, preload_q, comp_q, log_level=logging.INFO, cpu_priority=10, io_priority=None, cpu_shm_mode=False
    ):
        """Entry point of a daemon worker."""
        PersistentAsyncCaller.async_process_target(
            rank, queue, preload_q, comp_q, log_level, cpu_priority, io_priority, cpu_shm_mode
        )


class _ActiveAsyncRequest(NamedTuple):
    """A scheduled call: its index, the caller that owns its writer, and the (frozen) request."""

    idx: int
    async_caller: AsyncCaller
    async_request: AsyncRequest


class AsyncCallsQueue(metaclass=ObjectTracker):
    """FIFO of in-flight async saves: ``schedule_async_request`` starts one, ``maybe_finalize_async_calls``
    finalizes those that finished (in order).  All methods are collective over the ranks."""

    _warmup_persistent_caller: Optional[PersistentAsyncCaller] = None  # pre-started worker, consumed once

    def __init__(
        self,
        persistent: bool = True,
        is_daemon: bool = True,
        cpu_priority: int = 10,
        io_priority: Optional[int] = None,
        sigterm_timeout: float = 30.0,
        cpu_shm_mode: bool = False,
    ):
        self.async_calls: deque[_ActiveAsyncRequest] = deque([])
        self.call_idx: int = -1
        self.persistent: bool = persistent
        self.is_daemon: bool = is_daemon
        self.cpu_priority = cpu_priority
        self.io_priority = io_priority
        self.sigterm_timeout = sigterm_timeout
        self.cpu_shm_mode = cpu_shm_mode
        self.persistent_caller: AsyncCaller = None

    def _new_persistent_caller(self) -> PersistentAsyncCaller:
        return PersistentAsyncCaller(
            is_daemon=self.is_daemon,
            cpu_priority=self.cpu_priority,
            io_priority=self.io_priority,
            sigterm_timeout=self.sigterm_timeout,
            cpu_shm_mode=self.cpu_shm_mode,
        )

    def _get_async_caller(self):
        if not self.persistent:
            return TemporalAsyncCaller()
        if self.persistent_caller is None:
            warmed = AsyncCallsQueue._warmup_persistent_caller
            if warmed is not None:
                AsyncCallsQueue._warmup_persistent_caller = None
                if warmed.process is not None and not warmed.process.is_alive():
                    logger.warning(
                        "Pre-warmed async caller process (PID %s) is no longer alive; starting a fresh worker.",
                        warmed.process.pid,
                    )
                    warmed.process.join()
                    warmed.process = None
                self.persistent_caller = warmed
            else:
                self.persistent_caller = self._new_persistent_caller()
        return self.persistent_caller

    @classmethod
    def warmup_persistent_caller(
        cls,
        rank: int,
        is_daemon: bool = True,
        cpu_priority: int = 10,
        io_priority: Optional[int] = None,
        sigterm_timeout: float = 30.0,
        cpu_shm_mode: bool = False,
    ):
        """Spawn the persistent worker ahead of the first checkpoint (hides ~seconds of interpreter start)."""
        if cls._warmup_persistent_caller is None:
            caller = PersistentAsyncCaller(
                is_daemon=is_daemon, cpu_priority=cpu_priority, io_priority=io_priority,
                sigterm_timeout=sigterm_timeout, cpu_shm_mode=cpu_shm_mode,
            )
            caller._start_worker(rank)
            caller.rank = rank
            cls._warmup_persistent_caller = caller

    def schedule_async_request(self, async_request: AsyncRequest) -> int:
        """Start ``async_request``; returns its ``call_idx``."""
        self.call_idx += 1
        caller = self._get_async_caller()
        if len(async_request._fields) != len(AsyncRequest._fields):
            # requests built against an older AsyncRequest definition
            async_request = AsyncRequest(**async_request._asdict())
        async_request = async_request.freeze()
        # finalize functions stay on the trainer: they are closures over managers / process groups
        caller.schedule_async_call(async_request._replace(call_idx=self.call_idx, finalize_fns=[]))
        self.async_calls.append(_ActiveAsyncRequest(self.call_idx, caller, async_request))
        return self.call_idx

    def maybe_finalize_async_calls(self, blocking=False, no_dist=False) -> List[int]:
        """Finalize, oldest first, every call that is done (waiting for all of them if ``blocking``); stops at
        the first one still running.  Returns the finalized ``call_idx`` list."""
        finalized = []
        while self.async_calls:
            head = self.async_calls[0]
            if not head.async_caller.is_current_async_call_done(blocking, no_dist):
                break
            with debug_time("finalize", logger):
                self.async_calls.popleft()
                finalized.append(
                    head.async_request._replace(call_idx=head.idx).execute_finalize_fns(
                        validate_matching_call_idx=(not no_dist)
                    )
                )
        return finalized

    def get_num_unfinalized_calls(self):
        return len(self.async_calls)

    def close(self, abort=False):
        """Finalize what is pending (unless aborting) and stop the persistent worker."""
        if not abort and (self.persistent is False or self.persistent_caller is not None):
            self.maybe_finalize_async_calls(blocking=True)
        if abort:
            for call in self.async_calls:
                if isinstance(call.async_caller, TemporalAsyncCaller):
                    call.async_caller.close(abort=True)
            self.async_calls.clear()
        if self.persistent and self.persistent_caller:
            self.persistent_caller.close(abort=abort)
        if AsyncCallsQueue._warmup_persistent_caller is not None:
            AsyncCallsQueue._warmup_persistent_caller.close(abort=abort)
            AsyncCallsQueue._warmup_persistent_caller = None
        self.call_idx = -1
        self.persistent_caller = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def abort_nvrx_checkpoint():
    """Abort every live ``AsyncCallsQueue`` (used by in-process restart, reference ``inprocess/abort.py:194-202``);
    the next save starts with a fresh worker."""
    for queue in AsyncCallsQueue.get_instances():
        queue.close(abort=True)
