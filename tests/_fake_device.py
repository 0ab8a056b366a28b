"""TEST INFRASTRUCTURE: a stand-in for the GPU at the C-ABI boundary, so that the engine's Python (which only runs with a
device) can be executed in the CPU-only build container.

Everything the engine does on the device goes through ``libnvrx_snap.so`` -- that is the design -- so replacing the ~20 device
entry points is enough: "device memory" is host memory, pack / scatter move bytes with memmove following the REAL planner's
layout (the planner never needed CUDA), the drain and the fill are memmoves, streams and events are tokens, the checksum kernel
is the oracle.  Host-side entry points (planner, host buffers, checksum chaining) stay the real library.  Tensors that should
look like CUDA tensors are host tensors wrapped in ``FakeCudaTensor``.

What this does NOT test: the kernels, stream ordering, pinning, real collectives.  It catches what a GPU box would otherwise
be needed for first: wrong attribute names, argument orders, offset arithmetic, and the hand-offs between the engine, the
writer process and the managers.  Used only by tests/test_engine_flow_cpu.py."""
import contextlib
import ctypes as C

import numpy as np
import torch

from oracle import crc_oracle as co


class FakeCudaTensor(torch.Tensor):
    """Host memory that answers ``is_cuda`` / ``device`` like a tensor on cuda:0."""

    is_cuda = property(lambda self: True)
    device = property(lambda self: torch.device("cuda", 0))

    def get_device(self):
        return 0

    @staticmethod
    def wrap(t: torch.Tensor) -> "FakeCudaTensor":
        return torch.Tensor._make_subclass(FakeCudaTensor, t.detach())

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **(kwargs or {}))
        if isinstance(out, torch.Tensor) and not isinstance(out, FakeCudaTensor) and func.__name__ in ("detach", "contiguous", "clone", "view", "reshape", "__getitem__", "zero_", "fill_", "add_"):
            return FakeCudaTensor.wrap(out) if func.__name__ != "clone" else FakeCudaTensor.wrap(out)
        return out


def plain(t: torch.Tensor) -> torch.Tensor:
    """The same memory as an ordinary host tensor (for comparisons)."""
    return t.as_subclass(torch.Tensor) if isinstance(t, FakeCudaTensor) else t


def _val(x):
    return x.value if hasattr(x, "value") else (x or 0)


class FakeDeviceLib:
    def __init__(self, real):
        self._real = real
        self._mem = {}     # ptr -> numpy array ("device" allocations)
        self._plans = {}   # plan handle -> dict(ptrs, nbytes, flags)
        self._crcs = {}    # crc handle -> (offsets, nbytes)
        self._next = 100
        self.calls = []

    def __getattr__(self, name):  # planner, host buffers, checksum chaining, strerror ...: the real library
        return getattr(self._real, name)

    # ---- device memory, streams, events ---------------------------------------------------------------------------
    def nvrx_dev_alloc(self, device, nbytes, out):
        arr = np.zeros(max(int(nbytes), 512), dtype=np.uint8)
        self._mem[arr.ctypes.data] = arr
        out._obj.value = arr.ctypes.data
        return 0

    def nvrx_dev_free(self, device, ptr):
        self._mem.pop(_val(ptr), None)
        return 0

    def _token(self, out):
        self._next += 1
        out._obj.value = self._next
        return 0

    def nvrx_stream_create(self, device, prio, out):
        return self._token(out)

    def nvrx_event_create(self, device, timing, out):
        return self._token(out)

    def nvrx_stream_destroy(self, h):
        return 0

    nvrx_event_destroy = nvrx_stream_sync = nvrx_event_sync = nvrx_stream_destroy

    def nvrx_stream_wait_event(self, s, e):
        return 0

    nvrx_event_record = nvrx_plan_commit = nvrx_stream_wait_event

    def nvrx_event_query(self, e, out):
        out._obj.value = 1
        return 0

    def nvrx_event_elapsed_ms(self, a, b, out):
        out._obj.value = 1.0
        return 0

    def nvrx_hostbuf_create(self, name, nbytes, threads, pin, device, out):
        return self._real.nvrx_hostbuf_create(name, nbytes, threads, 0, device, out)  # nothing to pin

    # ---- planner bookkeeping (the real planner does the layout; remember what the walkers would read) ---------------
    def nvrx_plan_create_at(self, n, ptrs, nbytes, flags, offs, align, tile, device, out):
        rc = self._real.nvrx_plan_create_at(n, ptrs, nbytes, flags, offs, align, tile, device, out)
        if rc == 0:
            self._plans[out._obj.value] = {
                "ptrs": [ptrs[i] or 0 for i in range(n)], "nbytes": [nbytes[i] for i in range(n)],
                "flags": [flags[i] if flags else 0 for i in range(n)],
            }
        return rc

    def nvrx_plan_update_ptrs(self, h, ptrs):
        rc = self._real.nvrx_plan_update_ptrs(h, ptrs)
        if rc == 0:
            p = self._plans[_val(h)]
            p["ptrs"] = [ptrs[i] or 0 for i in range(len(p["nbytes"]))]
        return rc

    def _layout(self, h):
        p = self._plans[_val(h)]
        n = len(p["nbytes"])
        offs, packed = (C.c_uint64 * max(n, 1))(), (C.c_uint64 * max(n, 1))()
        assert self._real.nvrx_plan_layout(h, offs, packed) == 0
        return p, [offs[i] for i in range(n)], [packed[i] for i in range(n)]

    # ---- the walkers ------------------------------------------------------------------------------------------------
    def nvrx_pack(self, h, staging, stream):
        p, offs, packed = self._layout(h)
        for ptr, nb, fl, off, pk in zip(p["ptrs"], p["nbytes"], p["flags"], offs, packed):
            if not nb:
                continue
            if fl & 1:  # fp32 -> bf16 (RNE), as the narrow walker does
                src = torch.frombuffer((C.c_uint8 * nb).from_address(ptr), dtype=torch.float32)
                dst = torch.frombuffer((C.c_uint8 * pk).from_address(_val(staging) + off), dtype=torch.bfloat16)
                dst.copy_(src.to(torch.bfloat16))
            else:
                C.memmove(_val(staging) + off, ptr, nb)
        self.calls.append("pack")
        return 0

    def nvrx_scatter(self, h, staging, stream):
        p, offs, packed = self._layout(h)
        for ptr, nb, fl, off, pk in zip(p["ptrs"], p["nbytes"], p["flags"], offs, packed):
            if not nb:
                continue
            if fl & 1:  # bf16 -> fp32 (exact)
                src = torch.frombuffer((C.c_uint8 * pk).from_address(_val(staging) + off), dtype=torch.bfloat16)
                dst = torch.frombuffer((C.c_uint8 * nb).from_address(ptr), dtype=torch.float32)
                dst.copy_(src.to(torch.float32))
            else:
                C.memmove(ptr, _val(staging) + off, nb)
        self.calls.append("scatter")
        return 0

    def nvrx_drain(self, host, staging, nbytes, chunk, progress, base, stream, done):
        C.memmove(_val(host), _val(staging), nbytes)
        if _val(progress):
            C.c_uint64.from_address(_val(progress)).value = base + nbytes
        self.calls.append("drain")
        return 0

    def nvrx_fill(self, staging, host, nbytes, chunk, stream, done):
        C.memmove(_val(staging), _val(host), nbytes)
        self.calls.append("fill")
        return 0

    def nvrx_fill_from_fd(self, staging, staging_bytes, fd, n, stg_offs, nbytes, file_offs, chunk, slots, threads, device, stream):
        import os

        for i in range(n):
            if nbytes[i]:
                data = os.pread(fd, nbytes[i], file_offs[i])
                assert len(data) == nbytes[i] and stg_offs[i] + nbytes[i] <= staging_bytes
                C.memmove(_val(staging) + stg_offs[i], data, nbytes[i])
        self.calls.append("fill_from_fd")
        return 0

    def nvrx_snapshot(self, h, staging, host, chunk, progress, base, pack_stream, side, packed_ev, done):
        self.nvrx_pack(h, staging, pack_stream)
        total = C.c_uint64()
        assert self._real.nvrx_plan_info(h, C.byref(total), None, None) == 0
        return self.nvrx_drain(host, staging, total.value, chunk, progress, base, side, done)

    # ---- checksums --------------------------------------------------------------------------------------------------
    def nvrx_crc_create(self, n, offsets, nbytes, device, out):
        rc = self._real.nvrx_crc_create(n, offsets, nbytes, device, out)
        if rc == 0:
            self._crcs[out._obj.value] = ([offsets[i] for i in range(n)], [nbytes[i] for i in range(n)])
        return rc

    def nvrx_crc_run(self, h, dev_base, host_values, host_ready, ready_value, stream):
        offsets, nbytes = self._crcs[_val(h)]
        base = _val(dev_base)
        vals = [co.chunk_value(C.string_at(base + off, rows * 512)) for off, rows, _ in co.chunks_of(offsets, nbytes)]
        if vals:
            C.memmove(_val(host_values), (C.c_uint32 * len(vals))(*vals), 4 * len(vals))
        if _val(host_ready):
            C.c_uint64.from_address(_val(host_ready)).value = ready_value
        self.calls.append("crc")
        return 0


@contextlib.contextmanager
def fake_device(monkeypatch):
    """Yield ``(engine, lib)``: a SnapshotEngine for "cuda:0" running on the fake device layer (registered as the singleton the
    package's code obtains through ``SnapshotEngine.get()``)."""
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi
    from nvidia_resiliency_ext.checkpointing.b200 import engine as eng

    lib = FakeDeviceLib(_cabi.lib())
    monkeypatch.setattr(_cabi, "_LIB", lib)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(eng.SnapshotEngine, "_current_stream", lambda self: 0)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self if isinstance(self, FakeCudaTensor) else FakeCudaTensor.wrap(self))
    real_empty = torch.empty

    def empty(*args, **kwargs):
        dev = kwargs.get("device")
        if dev is not None and torch.device(dev).type == "cuda":
            kwargs["device"] = "cpu"
            return FakeCudaTensor.wrap(real_empty(*args, **kwargs))
        return real_empty(*args, **kwargs)

    monkeypatch.setattr(torch, "empty", empty)
    monkeypatch.setattr(eng.SnapshotEngine, "_instances", {})
    engine = eng.SnapshotEngine.get(0)
    try:
        yield engine, lib
    finally:
        eng.SnapshotEngine.shutdown_all()
