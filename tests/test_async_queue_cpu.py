"""Host logic of the async-call queue on CPU (world of one rank, gloo): scheduling order, finalize functions,
persistent worker life cycle, abort, worker-death detection.  Mirrors what the reference checks in
tests/checkpointing/unit/test_async_save.py, minus the GPU."""
import os
import time

import pytest
import torch

from nvidia_resiliency_ext.checkpointing.async_ckpt.core import (
    AsyncCallsQueue,
    AsyncRequest,
    PersistentAsyncCaller,
    TemporalAsyncCaller,
    abort_nvrx_checkpoint,
)


def _write(obj, path, delay=0.0):
    if delay:
        time.sleep(delay)
    torch.save(obj, path)


def _die(obj, path):
    os._exit(3)


def test_async_request_contract():
    calls = []
    req = AsyncRequest(None, (), [lambda: calls.append("a")])
    req.add_finalize_fn(lambda: calls.append("b"))
    frozen = req.freeze()
    assert frozen.is_frozen and not req.is_frozen
    with pytest.raises(RuntimeError):
        frozen.add_finalize_fn(lambda: None)
    assert frozen._replace(call_idx=4).execute_finalize_fns(validate_matching_call_idx=False) == 4
    assert calls == ["a", "b"]
    assert AsyncRequest._fields == ("async_fn", "async_fn_args", "finalize_fns", "async_fn_kwargs", "preload_fn", "is_frozen", "call_idx")


def test_execute_sync_uses_preload_as_payload(tmp_path, dist_1rank):
    out = tmp_path / "x.pt"
    done = []
    req = AsyncRequest(_write_args_swapped, (out, None), [lambda: done.append(1)], {}, preload_fn=lambda: {"v": torch.arange(3)})
    req.execute_sync()
    assert torch.equal(torch.load(out)["v"], torch.arange(3)) and done == [1]


def _write_args_swapped(path, payload):
    torch.save(payload, path)


@pytest.mark.parametrize("persistent", [True, False])
def test_schedule_finalize_order(tmp_path, dist_1rank, persistent):
    q = AsyncCallsQueue(persistent=persistent, cpu_shm_mode=True)
    order = []
    idxs = []
    for i in range(3):
        sd = {"i": torch.full((4,), i)}
        req = AsyncRequest(_write, (sd, tmp_path / f"f{i}.pt"), [lambda i=i: order.append(i)], {"delay": 0.05})
        idxs.append(q.schedule_async_request(req))
    assert idxs == [0, 1, 2] and q.get_num_unfinalized_calls() == 3
    finalized = q.maybe_finalize_async_calls(blocking=True, no_dist=True)
    assert finalized == [0, 1, 2] and order == [0, 1, 2] and q.get_num_unfinalized_calls() == 0
    for i in range(3):
        assert torch.equal(torch.load(tmp_path / f"f{i}.pt")["i"], torch.full((4,), i))
    q.close()
    assert q.call_idx == -1


def test_nonblocking_finalize_returns_immediately(tmp_path, dist_1rank):
    q = AsyncCallsQueue(persistent=True, cpu_shm_mode=True)
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.zeros(2)}, tmp_path / "w.pt"), [], {}))  # warm-up (spawn)
    q.maybe_finalize_async_calls(blocking=True, no_dist=True)
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.zeros(2)}, tmp_path / "s.pt"), [], {"delay": 0.5}))
    t0 = time.time()
    assert q.maybe_finalize_async_calls(blocking=False, no_dist=True) == []
    assert time.time() - t0 < 0.4
    t0 = time.time()
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=True) == [1]
    # completion is noticed within the poll slice, not the reference's 100 ms sleep quantum
    assert 0.2 < time.time() - t0 < 3.0  # (generous upper bound: CI machines stall)
    q.close()


def test_persistent_worker_lifecycle_and_abort(tmp_path, dist_1rank):
    q = AsyncCallsQueue(persistent=True, cpu_shm_mode=True)
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.ones(2)}, tmp_path / "a.pt"), [], {}))
    q.maybe_finalize_async_calls(blocking=True, no_dist=True)
    caller = q._get_async_caller()
    assert isinstance(caller, PersistentAsyncCaller) and caller._debug_is_async_process_running()
    assert q in AsyncCallsQueue.get_instances()
    abort_nvrx_checkpoint()
    assert not caller._debug_is_async_process_running() and q.persistent_caller is None
    # a fresh worker is started by the next save
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.ones(2)}, tmp_path / "b.pt"), [], {}))
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=True) == [0]
    assert (tmp_path / "b.pt").exists()
    q.close()


def test_dead_persistent_worker_is_detected(tmp_path, dist_1rank):
    """The reference spins forever here (core.py:577-589); the rebuilt queue raises."""
    q = AsyncCallsQueue(persistent=True, cpu_shm_mode=True)
    q.schedule_async_request(AsyncRequest(_die, ({}, tmp_path / "never.pt"), [], {}))
    with pytest.raises(RuntimeError, match="worker died"):
        q.maybe_finalize_async_calls(blocking=True, no_dist=True)
    q.close(abort=True)


def test_temporal_caller_forks_and_joins(tmp_path, dist_1rank):
    q = AsyncCallsQueue(persistent=False)
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.arange(5)}, tmp_path / "t.pt"), [], {"delay": 0.2}))
    caller = q.async_calls[0].async_caller
    assert isinstance(caller, TemporalAsyncCaller) and caller._debug_is_async_process_running()
    assert q.maybe_finalize_async_calls(blocking=False, no_dist=True) == []
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=True) == [0]
    assert not caller._debug_is_async_process_running()
    assert torch.equal(torch.load(tmp_path / "t.pt")["a"], torch.arange(5))
    q.close()


def test_warmup_persistent_caller(tmp_path, dist_1rank):
    AsyncCallsQueue.warmup_persistent_caller(0, cpu_shm_mode=True)
    warmed = AsyncCallsQueue._warmup_persistent_caller
    assert warmed is not None and warmed._debug_is_async_process_running()
    q = AsyncCallsQueue(persistent=True, cpu_shm_mode=True)
    assert q._get_async_caller() is warmed and AsyncCallsQueue._warmup_persistent_caller is None
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.ones(1)}, tmp_path / "w.pt"), [], {}))
    q.maybe_finalize_async_calls(blocking=True, no_dist=True)
    q.close()


def test_dist_finalize_validates_call_idx(tmp_path, dist_1rank):
    q = AsyncCallsQueue(persistent=False)
    q.schedule_async_request(AsyncRequest(_write, ({"a": torch.ones(1)}, tmp_path / "d.pt"), [], {}))
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=False) == [0]
    q.close()


def test_debug_time_log_format(caplog):
    import logging

    from nvidia_resiliency_ext.checkpointing.utils import debug_time

    log = logging.getLogger("nvrx_test")
    with caplog.at_level(logging.DEBUG, logger="nvrx_test"):
        with debug_time("outer", log):
            with debug_time("finalize_fn"):
                pass
    msgs = [r.getMessage() for r in caplog.records]
    assert any(m.startswith("outer.finalize_fn took ") and m.endswith("s") for m in msgs)
    assert any(m.startswith("outer took ") for m in msgs)


def test_torch_async_checkpoint_with_host_state_dict(tmp_path, dist_1rank):
    """BASELINE config C1 through TorchAsyncCheckpoint itself: a state dict without CUDA tensors needs no engine (and the
    persistent worker must come up on a machine without a GPU -- the reference's worker divides by device_count() there)."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint

    g = torch.Generator().manual_seed(0)
    sd = {f"p{i}": torch.randn(256, 64, generator=g) for i in range(8)}
    sd["meta"] = {"step": 7, "ids": torch.arange(5)}
    for persistent in (True, False):
        ckpt = TorchAsyncCheckpoint(persistent_queue=persistent)
        out = tmp_path / f"c1_{persistent}.pt"
        ckpt.async_save(sd, out)
        ckpt.save(sd, tmp_path / "sync.pt")
        ckpt.finalize_async_save(blocking=True)
        a, s = torch.load(out), torch.load(tmp_path / "sync.pt")
        assert list(a) == list(s) and all(torch.equal(a[k], s[k]) for k in sd if k != "meta")
        assert a["meta"]["step"] == 7 and torch.equal(a["meta"]["ids"], torch.arange(5))
        assert ckpt._get_async_calls_queue().get_num_unfinalized_calls() == 0
        ckpt.close()
