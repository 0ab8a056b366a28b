"""GPU tests of the reference-facing API (TorchAsyncCheckpoint, LocalCheckpointManager, BasicTensorAwareStateDict)
on top of the engine; structure follows reference tests/checkpointing/unit/test_async_save.py and test_basic_local.py.
The oracle's ``reference_preload`` / ``reference_snapshot_file`` provide the reference's snapshot of the same state."""
import os
import time

import pytest
import torch

from conftest import GOLDEN
from oracle import snapshot_oracle as orc

pytestmark = pytest.mark.gpu


def bit_equal(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    return a.numel() == 0 or torch.equal(a.contiguous().view(-1).view(torch.uint8), b.contiguous().view(-1).view(torch.uint8))


def model_state(ntensor=10, size=(1024, 1024), seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {f"param_{i}": torch.rand(size, device="cuda", generator=g) for i in range(ntensor)}
    sd["nested"] = {"ints": torch.randint(0, 100, (1000,), device="cuda", generator=g), "list": [torch.tensor(3.0, device="cuda"), "text", 7]}
    sd["half"] = torch.rand(333, device="cuda", generator=g).to(torch.bfloat16)
    return sd


def assert_tree_equal(x, y):
    if isinstance(x, dict):
        assert list(x) == list(y)
        for k in x:
            assert_tree_equal(x[k], y[k])
    elif isinstance(x, list):
        assert len(x) == len(y)
        for a, b in zip(x, y):
            assert_tree_equal(a, b)
    elif isinstance(x, torch.Tensor):
        assert bit_equal(x, y)
    else:
        assert x == y


@pytest.mark.parametrize("persistent", [True, False])
def test_async_is_equivalent_to_sync_and_to_reference(shm_dir, dist_1rank, built_library, persistent):
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint

    sd = model_state()
    ckpt = TorchAsyncCheckpoint(persistent_queue=persistent)
    ckpt.async_save(sd, shm_dir / "async.pt")
    ckpt.save(sd, shm_dir / "sync.pt")
    orc.reference_snapshot_file(sd, shm_dir / "reference.pt")  # what the reference writes for this state dict
    ckpt.finalize_async_save(blocking=True)
    a = torch.load(shm_dir / "async.pt", map_location="cuda", weights_only=False)
    s = torch.load(shm_dir / "sync.pt", map_location="cuda", weights_only=False)
    r = torch.load(shm_dir / "reference.pt", weights_only=False)
    assert_tree_equal(s, a)
    assert_tree_equal(sd, a)
    assert_tree_equal(r, torch.load(shm_dir / "async.pt", weights_only=False))
    ckpt.close()


def test_async_save_does_not_stall_and_sees_a_consistent_snapshot(shm_dir, dist_1rank, built_library):
    """The snapshot is taken in stream order: mutating the parameters right after async_save must not leak into
    the file (the reference gets the same guarantee from its device-wide sync)."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint

    g = torch.Generator(device="cuda").manual_seed(1)
    sd = {f"p{i}": torch.rand(4096, 4096, device="cuda", generator=g) for i in range(16)}  # 1 GiB
    want = {k: v.clone() for k, v in sd.items()}
    ckpt = TorchAsyncCheckpoint()
    ckpt.async_save({"a": torch.ones(8, device="cuda")}, shm_dir / "warm.pt")  # spawn + pin outside the timing
    ckpt.finalize_async_save(blocking=True)
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine

    SnapshotEngine.get().reserve(sum(v.numel() * 4 for v in sd.values()) + (1 << 20))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ckpt.async_save(sd, shm_dir / "big.pt")
    dt = time.perf_counter() - t0
    for v in sd.values():  # "the next optimizer step"
        v.add_(1.0)
    ckpt.finalize_async_save(blocking=True)
    assert dt < 0.5, f"async_save blocked for {dt:.3f}s"  # 1 GiB over PCIe alone would be ~20 ms; the call only enqueues
    loaded = torch.load(shm_dir / "big.pt", weights_only=False)
    for k in want:
        assert bit_equal(loaded[k], want[k])
    ckpt.close()


def test_narrowed_async_save(shm_dir, dist_1rank, built_library):
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint

    sd = model_state(ntensor=3, size=(257, 129))
    sd["edge"] = torch.tensor([float("nan"), float("inf"), -0.0, 1e-40, 1.00390625], device="cuda")
    ckpt = TorchAsyncCheckpoint(narrow_fp32_to_bf16=True)
    ckpt.async_save(sd, shm_dir / "narrow.pt")
    ckpt.finalize_async_save(blocking=True)
    got = torch.load(shm_dir / "narrow.pt", weights_only=False)
    for (k, a), b in zip(((k, v) for k, v in got.items() if isinstance(v, torch.Tensor)), (v for v in sd.values() if isinstance(v, torch.Tensor))):
        want = b.to(torch.bfloat16) if b.dtype == torch.float32 else b
        assert bit_equal(a, want), k
    ckpt.close()


def test_persistent_worker_abort_and_resume(shm_dir, dist_1rank, built_library):
    """reference test_async_save.py:58 -- abort the persistent worker, then save again."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import abort_nvrx_checkpoint
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint

    sd = model_state(ntensor=2, size=(64, 64))
    ckpt = TorchAsyncCheckpoint(persistent_queue=True)
    ckpt.async_save(sd, shm_dir / "one.pt")
    ckpt.finalize_async_save(blocking=True)
    q = ckpt._get_async_calls_queue()
    caller = q._get_async_caller()
    assert caller._debug_is_async_process_running()
    abort_nvrx_checkpoint()
    assert not caller._debug_is_async_process_running()
    ckpt.async_save(sd, shm_dir / "two.pt")
    ckpt.finalize_async_save(blocking=True)
    assert_tree_equal(sd, torch.load(shm_dir / "two.pt", map_location="cuda", weights_only=False))
    ckpt.close()


# ---- TensorAwareStateDict + LocalCheckpointManager ---------------------------------------------------
def tasd(seed=0, n=50):
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict

    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {"model": {f"w{i}": torch.empty(128, 128, device="cuda").random_(generator=g) for i in range(n)},
          "opt": [{"step": torch.tensor(float(i), device="cuda"), "m": torch.randn(77, device="cuda", generator=g)} for i in range(3)],
          "iteration": seed}
    return BasicTensorAwareStateDict(sd)


def tasd_equal(a, b):
    ta, tb = list(a.tensors), list(b.tensors)
    return len(ta) == len(tb) and all(bit_equal(x, y) for x, y in zip(ta, tb)) and a.state_dict["iteration"] == b.state_dict["iteration"]


def test_tasd_contract(built_library):
    t = tasd(1, n=4)
    ref = orc.flatten_tensors(t.state_dict)
    assert [x.data_ptr() for x in t.tensors] == [x.data_ptr() for x in ref]  # same flattening order as the oracle
    popped = t.pop_tensors()
    assert t.is_hollow
    with pytest.raises(AssertionError):
        t.pop_tensors()
    import pickle

    assert len(pickle.dumps(t)) < 10_000  # hollow skeleton is tiny
    t.insert_tensors(popped)
    assert not t.is_hollow and [x.data_ptr() for x in t.tensors] == [x.data_ptr() for x in popped]
    t.pop_tensors()
    t.init_tensors()
    assert all(x.is_cuda and x.shape == y.shape and x.dtype == y.dtype for x, y in zip(t.tensors, popped))


def test_copy_to_cpu_matches_reference_snapshot_and_restores(built_library):
    t = tasd(2)
    want = [x.clone() for x in t.tensors]
    reference = orc.flatten_tensors(orc.reference_preload(t.state_dict))  # the reference's snapshot of the same state
    snap = t.copy_tensors_to_cpu(non_blocking=True)
    snap.wait()
    torch.cuda.synchronize()
    assert all(not x.is_cuda for x in t.tensors)
    assert all(bit_equal(a, b) for a, b in zip(t.tensors, reference))
    t.restore_tensor_device()
    assert all(x.is_cuda and bit_equal(x, w) for x, w in zip(t.tensors, want))
    snap.release()


@pytest.mark.parametrize("use_ramdisk", [True, False])
@pytest.mark.parametrize("is_async", [True, False])
def test_basic_save_load_scenarios(tmp_path, shm_dir, dist_1rank, built_library, use_ramdisk, is_async):
    """reference tests/checkpointing/unit/test_basic_local.py:47"""
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    root = (shm_dir if use_ramdisk else tmp_path) / "subdir"

    def save(mgr, sd, it):
        req = mgr.save(sd, it, is_async)
        if is_async:
            req.execute_sync()
        else:
            assert req is None

    mgr = LocalCheckpointManager(root)
    sd1 = tasd(1)
    save(mgr, sd1, 1)
    assert mgr.find_latest() == 1
    loaded, cid = mgr.load()
    sd1.restore_tensor_device()
    assert tasd_equal(loaded, sd1) and tasd_equal(loaded, tasd(1)) and cid == (1, 0, "")
    assert all(x.is_cuda for x in loaded.tensors)

    mgr = LocalCheckpointManager(root)
    assert mgr.find_latest() == 1
    loaded, cid = mgr.load()
    assert tasd_equal(loaded, tasd(1))

    mgr = LocalCheckpointManager(root)
    first = mgr._local_ckpt_path_from_id(mgr._ckpt_id(1))
    os.remove(first)
    assert mgr.find_latest() == -1
    save(mgr, tasd(1), 1)
    assert first.exists()
    save(mgr, tasd(2), 2)
    time.sleep(0.4)
    assert not first.exists() and mgr._local_ckpt_path_from_id(mgr._ckpt_id(2)).exists()


def test_async_local_save_through_fork_queue(shm_dir, dist_1rank, built_library):
    """The documented combination (examples/checkpointing/local_ckpt.py): AsyncCallsQueue(persistent=False); the forked
    writer follows the drain through shared memory while training mutates the parameters."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(shm_dir)
    q = AsyncCallsQueue(persistent=False)
    sd = tasd(7, n=200)
    live = list(sd.tensors)
    req = mgr.save(sd, 7, is_async=True)
    q.schedule_async_request(req)
    for x in live:  # next training step overwrites the parameters
        x.zero_()
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=False) == [0]
    assert mgr.find_latest() == 7
    loaded, _ = mgr.load()
    assert tasd_equal(loaded, tasd(7, n=200))
    q.close()


def test_engine_loads_the_references_own_snapshot_file(dist_1rank, built_library, tmp_path):
    """C5 bit-exact check against the reference's file: tests/golden/iter_0000007_0_local.pt was WRITTEN BY THE REFERENCE;
    load it through the product (H2D + scatter kernel) and compare with the inputs the reference was given."""
    import shutil

    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(tmp_path)
    mgr._ensure_dir()
    shutil.copy(GOLDEN / "iter_0000007_0_local.pt", mgr.local_ckpt_dir / "iter_0000007_0_local.pt")
    assert mgr.find_latest() == 7
    loaded, cid = mgr.load()
    assert cid == (7, 0, "")
    inputs = orc.flatten_tensors(torch.load(GOLDEN / "local_inputs.pt", weights_only=False))
    got = list(loaded.tensors)
    assert len(got) == len(inputs) and all(x.is_cuda and bit_equal(x, y) for x, y in zip(got, inputs))
    assert loaded.state_dict["rng"][1:] == ["not-a-tensor", 5]


def test_many_tensor_finalize_is_fast(shm_dir, dist_1rank, built_library, caplog):
    """reference tests/checkpointing/unit/test_cleanup.py: finalize_fn of an async local save stays in the tens of ms."""
    import logging

    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    mgr = LocalCheckpointManager(shm_dir)
    q = AsyncCallsQueue(persistent=False)
    with caplog.at_level(logging.DEBUG):
        for it in (1, 2):
            q.schedule_async_request(mgr.save(tasd(it, n=2048), it, is_async=True))  # 128 MiB in 2048 tensors
            q.maybe_finalize_async_calls(blocking=True, no_dist=False)
    took = [float(r.getMessage().split(" took ")[1].rstrip("s")) for r in caplog.records if "finalize_fn took" in r.getMessage()]
    assert len(took) == 2 and max(took) < 0.1, took
    q.close()


def test_mcore_shaped_state_dict_through_the_engine(shm_dir, dist_1rank, built_library):
    """Row f4: a third-party TensorAwareStateDict shaped like Megatron-Core's MCoreTensorAwareStateDict -- tensors
    inside ShardedTensor-like objects, a `common` part with a host tensor, its own per-tensor device<->host methods -- takes
    the engine path through the ABC contract alone (pop -> ONE pack + drain -> insert; restore: file -> ring -> scatter)."""
    from _mcore_like import MCoreLikeTensorAwareStateDict

    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue
    from nvidia_resiliency_ext.checkpointing.b200.engine import SnapshotEngine
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    g = torch.Generator(device="cuda").manual_seed(11)
    model = {f"layers.{i}.w": torch.randn(257 + i, 129, device="cuda", generator=g) for i in range(6)}
    optim = {i: {"exp_avg": torch.randn(257 + i, 129, device="cuda", generator=g), "step": torch.tensor(float(i), device="cuda")} for i in range(6)}
    tasd_ = MCoreLikeTensorAwareStateDict.from_state_dict(model, optim, iteration=9)
    want = [t.clone() for t in tasd_.tensors]
    engine = SnapshotEngine.get()
    launches, restores = engine.launches, engine.file_restores
    mgr = LocalCheckpointManager(shm_dir / "mcore")
    q = AsyncCallsQueue(persistent=False)
    try:
        req = mgr.save(tasd_, 3, is_async=True)
        assert tasd_.calls == [] and engine.launches > launches  # the engine packed it; the class's own copy loop did not run
        for t in model.values():
            t.zero_()  # training goes on: the snapshot was consistent at save() in stream order
        q.schedule_async_request(req)
        q.maybe_finalize_async_calls(blocking=True, no_dist=False)
        mgr2 = LocalCheckpointManager(shm_dir / "mcore")
        assert mgr2.find_latest() == 3
        loaded, cid = mgr2.load()
        assert cid == (3, 0, "") and loaded.calls == [] and engine.file_restores == restores + 1
        assert loaded.common["iteration"] == 9 and torch.equal(loaded.common["rng_state"], torch.arange(16, dtype=torch.uint8))
        got = list(loaded.tensors)
        assert len(got) == len(want) and all(a.is_cuda and bit_equal(a, b) for a, b in zip(got, want))
        # and what the reference implementation of the same flow (the class's own methods) stores is the same bytes
        os.environ["NVRX_B200_GENERIC_TASD"] = "0"
        try:
            ref = MCoreLikeTensorAwareStateDict.from_state_dict({k: w.clone() for k, w in zip(model, want[:6])}, {}, iteration=9)
            mgr3 = LocalCheckpointManager(shm_dir / "mcore_ref")
            mgr3.save(ref, 1, is_async=False)
            assert ref.calls == ["copy_tensors_to_cpu"]
            assert mgr3.find_latest() == 1
            back, _ = mgr3.load()
            assert back.calls[-1] == "restore_tensor_device"  # (the list is pickled with the object: it also holds the save-side call)
            assert all(bit_equal(a, b) for a, b in zip(back.tensors, want[:6]))
        finally:
            os.environ.pop("NVRX_B200_GENERIC_TASD", None)
    finally:
        q.close()
