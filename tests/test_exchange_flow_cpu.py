"""Control flow and byte movement of the two exchange variants that were written without GPU time, exercised on the CPU with
a stand-in engine: "device" pointers are host pointers, the pack is the oracle's pack, nvrx_drain is a memmove.  Clique of one
(no collective is issued), so what is checked is everything around the collective: planning, ring / slice arithmetic, the
per-chunk drains, progress accounting, the host views and the Snapshot handles."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import snapshot_oracle as orc


class _Ev:
    def __init__(self, *a, **k):
        self.handle = 1

    def record(self, stream):
        pass

    def synchronize(self):
        pass


class _Stream:
    handle = 7

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        assert isinstance(ev, _Ev)

    def synchronize(self):
        pass


class _Lib:
    """nvrx_drain(host_dst, staging, bytes, chunk, progress, base, stream, done_event) on host memory."""

    def __init__(self):
        self.drains = []

    def nvrx_drain(self, dst, src, nbytes, chunk, progress, base, stream, done):
        C.memmove(dst, src, nbytes)
        C.c_uint64.from_address(progress).value = base + nbytes
        self.drains.append((dst, src, nbytes, base))
        return 0


class _Staging:
    def __init__(self, nbytes):
        self.arr = np.zeros(max(nbytes, 512), dtype=np.uint8)
        self.ptr = self.arr.ctypes.data
        self.nbytes = self.arr.size


class _DevBuf(_Staging):
    """DeviceBuffer stand-in (host memory)."""

    def __init__(self, nbytes, device=0):
        super().__init__(nbytes)

    def close(self):
        pass


class _FakeEngine:
    """The attributes of SnapshotEngine that b200/exchange.py touches."""

    device = 0
    align = 512
    drain_chunk = 256 << 20

    def __init__(self):
        from nvidia_resiliency_ext.checkpointing.b200.engine import _Slot

        self.lib = _Lib()
        self._staging_free = None
        self._side = _Stream()
        self._slots = [_Slot(i) for i in range(2)]
        self.max_host_slots = 4
        self.launches = 0
        self._staging = None
        self._n = 0

    def _current_stream(self):
        return 0

    def _ensure_staging(self, nbytes):
        if self._staging is None or self._staging.nbytes < nbytes:
            self._staging = _Staging(nbytes)
        return self._staging

    def _spare_slots(self):
        from nvidia_resiliency_ext.checkpointing.b200.engine import spare_slots

        return spare_slots(self._slots, self.max_host_slots)

    def _plan_for(self, tensors, narrow, container=False, offsets=None):
        from nvidia_resiliency_ext.checkpointing.b200 import ptzip
        from nvidia_resiliency_ext.checkpointing.b200.engine import Plan

        sizes = [t.numel() * t.element_size() for t in tensors]
        offs = ptzip.slot_offsets(sizes)[0] if container else None
        real = Plan([t.data_ptr() if n else 0 for t, n in zip(tensors, sizes)], sizes, None, device=0, staging_offsets=offs)

        class P:  # the real planner's layout, a host-side pack
            offsets, packed_nbytes, staging_bytes, n_tiles = real.offsets, real.packed_nbytes, real.staging_bytes, real.n_tiles

            @staticmethod
            def pack(staging_ptr, stream):
                for t, off, n in zip(tensors, real.offsets, sizes):
                    if n:
                        C.memmove(staging_ptr + off, t.data_ptr(), n)

        return P

    def _acquire_slot(self, nbytes, slot=None):
        from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer, choose_slot, _Slot

        slot, _ = choose_slot(self._slots, nbytes, self.max_host_slots)
        if slot is None:
            slot = _Slot(len(self._slots))
            self._slots.append(slot)
        if slot.buf is None or slot.buf.capacity < nbytes:
            if slot.buf is not None:
                slot.buf.close()
            self._n += 1
            slot.buf = HostBuffer.create(nbytes, name=f"/nvrx_fake_{os.getpid()}_{self._n}", pin=False, prefault_threads=1)
            slot.drained_total = 0
        slot.done_event = slot.done_event or _Ev()
        slot.busy = True
        return slot

    def close(self):
        for s in self._slots:
            if s.buf is not None:
                s.buf.close()


class _Group:
    my_group_rank = 0
    group = None

    def __init__(self, world=1, me=0):
        self.world_size = world
        self.my_group_rank = me

    def all_gather_int(self, v):
        return [v] * self.world_size

    def all_gather_object(self, obj):
        return [obj] * self.world_size


def _host_bytes(ptr, nbytes, device):
    """as_uint8_tensor stand-in: a uint8 tensor over host memory at ``ptr``."""
    return torch.frombuffer((C.c_uint8 * max(nbytes, 1)).from_address(ptr), dtype=torch.uint8)[:nbytes]


class _Dist:
    """all_gather_into_tensor stand-in for a clique whose members all hold what this member holds."""

    calls = 0

    @classmethod
    def all_gather_into_tensor(cls, out, inp, group=None):
        world = out.numel() // inp.numel()
        for r in range(world):
            out[r * inp.numel() : (r + 1) * inp.numel()] = inp
        cls.calls += 1


def _tensors():
    g = torch.Generator().manual_seed(3)
    return [torch.randn(700, 33, generator=g), torch.arange(11), torch.empty(0), torch.randn(4097, generator=g).to(torch.bfloat16),
            torch.tensor(5.0), torch.randint(0, 255, (300_001,), dtype=torch.uint8, generator=g)]


@pytest.fixture
def fake(monkeypatch, built_library):
    from nvidia_resiliency_ext.checkpointing.b200 import engine as eng_mod
    from nvidia_resiliency_ext.checkpointing.b200 import exchange as xch

    monkeypatch.setattr(xch, "Event", _Ev)
    monkeypatch.setattr(xch, "DeviceBuffer", _DevBuf)
    monkeypatch.setattr(eng_mod, "Stream", _Stream)
    monkeypatch.setattr(eng_mod, "stream_wait_event", lambda stream, ev: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: __import__("contextlib").nullcontext())
    monkeypatch.setattr(xch, "as_uint8_tensor", _host_bytes)
    monkeypatch.setattr(xch, "dist", _Dist)
    engine = _FakeEngine()
    yield engine, xch
    engine.close()


def _placeholders(tensors, world=1):
    from nvidia_resiliency_ext.checkpointing.local.replication.torch_device_utils import TensorPlaceholder

    return [[TensorPlaceholder(t) for t in tensors] for _ in range(world)]


def _bit_equal(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and (a.numel() == 0 or torch.equal(a.contiguous().view(-1).view(torch.uint8), b.contiguous().view(-1).view(torch.uint8)))


@pytest.mark.parametrize("chunk_kb", [1, 64, 1 << 20])
def test_streamed_exchange_moves_every_byte_once(fake, monkeypatch, chunk_kb):
    engine, xch = fake
    monkeypatch.setattr(xch, "STREAM_CHUNK", chunk_kb << 10)
    tensors = _tensors()
    geo = xch._geometry(_Group(), _placeholders(tensors), 512, 1)
    result, snaps = xch._allgather_streamed(engine, _Group(), tensors, geo, 1)
    assert engine.last_exchange == "nccl-streamed" and engine.launches == 1
    assert len(result) == 1 and all(_bit_equal(a, b) for a, b in zip(result[0], tensors))
    (snap,) = snaps
    slice_bytes = geo["slot_bytes"]
    assert snap.progress_target == slice_bytes and snap.slot.buf.progress == slice_bytes
    # the drains tile the slice in order, chunk by chunk, alternating source halves never overlapping in the destination
    pos = snap.slot.buf.data_ptr
    for dst, src, n, base in engine.lib.drains:
        assert dst == pos and base == pos - snap.slot.buf.data_ptr
        pos += n
    assert pos - snap.slot.buf.data_ptr == slice_bytes
    expect, _, _ = orc.pack_oracle(tensors)
    got = snap.slot.buf.as_tensor(len(expect)).numpy()
    offs, sizes = geo["layouts"][0][0], geo["layouts"][0][1]
    for o, n in zip(offs, sizes):
        assert np.array_equal(got[o : o + n], expect[o : o + n])
    # a second exchange of the same structure reuses geometry, slot views and continues the progress count
    snap.slot.busy = False
    views_before = [id(v) for v in result[0]]
    result2, (snap2,) = xch._allgather_streamed(engine, _Group(), tensors, geo, 1)
    assert snap2.slot is snap.slot and snap2.progress_target == 2 * slice_bytes
    assert [id(v) for v in result2[0]] == views_before


def test_container_landing_publishes_as_a_checkpoint(fake, monkeypatch, shm_dir):
    """Replicated zero-copy, host half: the member's slice lands in a slot of its own in container geometry and the writer
    publishes that slot as the member's file."""
    from nvidia_resiliency_ext.checkpointing.b200 import fastsave

    engine, xch = fake
    monkeypatch.setenv("NVRX_B200_ZERO_COPY", "1")
    monkeypatch.setenv("NVRX_B200_ZERO_COPY_REPLICAS", "1")
    tensors = _tensors()
    group = _Group()
    assert xch._container_exchange(engine, group, 1)
    geo = xch._geometry(group, _placeholders(tensors), 512, 1, container=True)
    plan = engine._plan_for(tensors, [False] * len(tensors), True)
    assert list(plan.offsets) == geo["layouts"][0][0]

    class X:  # the exchange buffer after the fused pack: this member's slice in container geometry
        pass

    xbuf = X()
    stg = _Staging(geo["slot_bytes"])
    plan.pack(stg.ptr, 0)
    xbuf.ptr = stg.ptr
    result, (snap,) = xch._land_per_member(engine, geo, xbuf, 1)
    assert all(_bit_equal(a, b) for a, b in zip(result[0], tensors))
    assert snap.progress_target == geo["layouts"][0][2] == snap.slot.buf.progress
    state = {"model": {"a": result[0][0], "b": result[0][1]}, "rest": result[0][2:], "it": 3}
    target = shm_dir / "iter_0000003_0_local.pt"
    desc = snap.descriptor()
    with fastsave.slot_ranges(fastsave.ranges_for([desc], [snap.slot.buf])):
        assert fastsave.save(state, str(target)) == "linked"
    assert os.path.samefile("/dev/shm" + snap.slot.buf.name, target)
    loaded = torch.load(target, weights_only=False)
    assert _bit_equal(loaded["model"]["a"], tensors[0]) and all(_bit_equal(a, b) for a, b in zip(loaded["rest"], tensors[2:]))
    # the published slot is not offered again while the file exists
    snap.slot.busy = False
    assert engine._spare_slots() == 1 + 2
    os.unlink(target)
    assert engine._spare_slots() == 2 + 2


@pytest.mark.parametrize("world,me", [(2, 0), (3, 2)])
def test_streamed_exchange_ring_arithmetic_for_several_members(fake, monkeypatch, world, me):
    """Same flow with a faked collective (every member contributes the same bytes): the two ring halves, the per-member source
    offsets inside a half and the per-member destination slices of the host slot."""
    engine, xch = fake
    monkeypatch.setattr(xch, "STREAM_CHUNK", 48 << 10)
    tensors = _tensors()
    group = _Group(world, me)
    geo = xch._geometry(group, _placeholders(tensors, world), 512, world)
    before = _Dist.calls
    result, (snap,) = xch._allgather_streamed(engine, group, tensors, geo, world)
    n_chunks = -(-geo["slot_bytes"] // (48 << 10))
    assert _Dist.calls - before == n_chunks and len(engine.lib.drains) == n_chunks * world
    assert snap.progress_target == world * geo["slot_bytes"] == snap.slot.buf.progress
    assert len(result) == world
    for r in range(world):
        assert all(_bit_equal(a, b) for a, b in zip(result[r], tensors)), r


@pytest.mark.parametrize("world,me", [(1, 0), (2, 1), (4, 2)])
def test_one_shot_exchange_with_cached_geometry(fake, monkeypatch, world, me):
    """The default replication path (pack into my slice + ONE collective + ONE drain), as refactored with the geometry and
    host-view caches: twice with the same structure, once with a changed one."""
    from nvidia_resiliency_ext.checkpointing.b200 import engine as eng_mod

    engine, xch = fake
    monkeypatch.setenv("NVRX_B200_EXCHANGE", "nccl")
    monkeypatch.setattr(eng_mod.SnapshotEngine, "get", classmethod(lambda cls, device=None, **kw: engine))
    tensors = _tensors()
    group = _Group(world, me)
    placeholders = _placeholders(tensors, world)

    monkeypatch.setattr(xch.SnapshotEngine, "get", classmethod(lambda cls, device=None, **kw: engine))
    runs = []
    for rep in range(2):
        result, (snap,) = xch.allgather_packed(group, tensors, placeholders, "cpu")
        assert engine.last_exchange == "nccl-allgather" and len(result) == world
        for r in range(world):
            assert all(_bit_equal(a, b) for a, b in zip(result[r], tensors)), (rep, r)
        assert snap.progress_target == (rep + 1) * world * group.__dict__["_packed_geometry"]["slot_bytes"]
        runs.append([id(v) for v in result[0]])
        snap.slot.busy = False
    assert runs[0] == runs[1]  # same slot buffer, same structure: the views are reused
    geo1 = group.__dict__["_packed_geometry"]
    # structure change (one tensor grows): new placeholder lists -> new geometry, new views, still correct
    tensors2 = tensors[:-1] + [torch.arange(12345, dtype=torch.int32)]
    result, (snap,) = xch.allgather_packed(group, tensors2, _placeholders(tensors2, world), "cpu")
    assert group.__dict__["_packed_geometry"] is not geo1
    for r in range(world):
        assert all(_bit_equal(a, b) for a, b in zip(result[r], tensors2)), r
