"""Randomised (hypothesis) checks of the host-side building blocks against their definitions: checksum chaining vs zlib,
container geometry invariants, and the planner's coverage in container geometry."""
import ctypes as C
import zlib

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import crc_oracle as co
from oracle import snapshot_oracle as orc

SETTINGS = dict(max_examples=int(__import__("os").environ.get("NVRX_TEST_EXAMPLES", "30")), deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])

sizes_strategy = st.lists(
    st.one_of(st.integers(0, 2048), st.integers(60_000, 70_000), st.integers(0, 300_000)), min_size=1, max_size=12
)


@settings(**SETTINGS)
@given(sizes=sizes_strategy, shifts=st.lists(st.sampled_from([0, 0, 0, 16, 8, 4]), min_size=12, max_size=12), seed=st.integers(0, 2**31))
def test_crc_finish_equals_zlib(built_library, sizes, shifts, seed):
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi

    offsets, cur = [], 0
    for nb, sh in zip(sizes, shifts):
        cur = -(-cur // 64) * 64 + sh
        offsets.append(cur)
        cur += nb
    payload = np.random.default_rng(seed).integers(0, 256, cur + 16, dtype=np.uint8)
    chunks = co.chunks_of(offsets, sizes)
    values = [co.chunk_value(payload[o : o + r * 512].tobytes()) for o, r, _ in chunks]
    n = len(sizes)
    out = (C.c_uint32 * n)()
    rc = _cabi.lib().nvrx_crc_finish(
        n, (C.c_uint64 * n)(*offsets), (C.c_uint64 * n)(*sizes), (C.c_uint32 * max(1, len(values)))(*values), len(values),
        payload.ctypes.data, out,
    )
    assert rc == 0
    assert list(out) == [zlib.crc32(payload[o : o + nb].tobytes()) for o, nb in zip(offsets, sizes)]


@settings(**SETTINGS)
@given(sizes=sizes_strategy)
def test_container_geometry_invariants(built_library, sizes):
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import Plan

    offsets, span = ptzip.slot_offsets(sizes)
    lay = ptzip.slot_layout(sizes)
    end = 0
    for i, (off, nb, rec) in enumerate(zip(offsets, sizes, lay.records)):
        assert off % 512 == 0 and off >= end
        # the gap in front of the data holds that record's local header: 30 bytes + name (+ zip64 sizes), padding >= 4 or 0
        header = rec.data_off - rec.header_off
        assert header >= 30 + len(rec.name) and rec.header_off == ptzip.SLOT_PREFIX + end
        end = off + nb
    assert span == end
    # the CUDA-free planner accepts the geometry as it is and covers every byte once
    ptrs = [0x10000000 + 0x100000 * i if nb else 0 for i, nb in enumerate(sizes)]
    plan = Plan(ptrs, sizes, None, device=0, staging_offsets=offsets)
    assert list(plan.offsets) == offsets and plan.staging_bytes >= span
    n_bulk, tiles = plan.tiles()
    covered = [0] * len(sizes)
    for seg, nb, _ in tiles:
        covered[seg] += nb
    assert covered == list(sizes)
    plan.close()


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(
    shapes=st.lists(st.tuples(st.sampled_from(["float32", "bfloat16", "int64", "uint8", "float64", "bool"]),
                              st.lists(st.integers(0, 40), min_size=0, max_size=3)), min_size=1, max_size=8),
    seed=st.integers(0, 2**31),
)
def test_published_slot_loads_back_for_any_mix_of_tensors(built_library, shm_dir, shapes, seed):
    import os

    from test_zero_copy_cpu import _same, _slot_with_snapshot
    from nvidia_resiliency_ext.checkpointing.b200.persist import save_snapshot_with_torch

    g = torch.Generator().manual_seed(seed)
    state = {"t": [], "meta": {"seed": seed}}
    for name, shp in shapes:
        dt = getattr(torch, name)
        base = torch.randint(0, 2 if dt == torch.bool else 100, tuple(shp), generator=g)
        state["t"].append(base.to(dt))
    os.environ["NVRX_B200_ZERO_COPY"] = "1"
    slot_name = f"/nvrx_prop_{os.getpid()}_{seed}"
    hb, desc, skeleton = _slot_with_snapshot(state, slot_name)
    target = shm_dir / f"p{seed}.pt"
    try:
        save_snapshot_with_torch(skeleton, str(target), desc)
        if any(t.numel() for t in orc.flatten_tensors(state)):
            assert os.path.samefile("/dev/shm" + slot_name, target)
        _same(torch.load(target, weights_only=False), state)
        _same(torch.load(target, weights_only=False, mmap=True), state)
    finally:
        os.environ.pop("NVRX_B200_ZERO_COPY", None)
        hb.close()
        if target.exists():
            target.unlink()


@settings(**SETTINGS)
@given(slot_kib=st.integers(1, 5000), limit=st.sampled_from([0, 512, 4096, 1 << 20, 256 << 20]), odd=st.integers(0, 1))
def test_stream_schedule_covers_the_slice_once(slot_kib, limit, odd):
    from nvidia_resiliency_ext.checkpointing.b200.exchange import stream_schedule

    slot_bytes = slot_kib * 1024 - odd * 512  # slices are multiples of 512 bytes
    chunk, steps = stream_schedule(slot_bytes, limit)
    assert chunk % 512 == 0 and chunk >= 512
    pos = 0
    for i, (lo, ln, half) in enumerate(steps):
        assert lo == pos and 0 < ln <= chunk and half == i % 2
        pos += ln
    assert pos == slot_bytes and all(ln == chunk for _, ln, _ in steps[:-1])


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(sizes=st.lists(st.one_of(st.integers(0, 64), st.integers(1, 5000), st.integers(60_000, 200_000)), min_size=1, max_size=9),
       host_first=st.booleans(), seed=st.integers(0, 2**31))
def test_parallel_writer_files_verify_and_restore_offsets_are_found(built_library, tmp_path_factory, sizes, host_first, seed):
    """Random state dicts through the default copying writer (skip_data container + slot writer + checksum patch): the file
    verifies like a torch.save file (zipfile.testzip), loads bit-exactly, and ``tensor_offsets_in_file`` finds where every
    mmap-loaded tensor lives -- also when a host tensor pickled first shifts the record numbering (Megatron-style `common`)."""
    import os
    import zipfile

    from nvidia_resiliency_ext.checkpointing.b200 import fastsave, ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer

    g = torch.Generator().manual_seed(seed)
    total = sum(-(-s // 512) * 512 + 512 for s in sizes) + 4096
    hb = HostBuffer.create(total, name=f"/nvrx_prop_{os.getpid()}_{seed % 100000}", pin=False, prefault_threads=1)
    try:
        views, off = [], 0
        for s in sizes:
            v = hb.segment(off, s, torch.uint8, (s,))
            if s:
                v.copy_(torch.randint(0, 256, (s,), dtype=torch.uint8, generator=g))
            views.append(v)
            off = -(-(off + s) // 512) * 512
        obj = {"slot": views, "meta": {"seed": seed}}
        if host_first:
            obj = {"common": {"rng": torch.arange(37, dtype=torch.int64)}, **obj}
        path = tmp_path_factory.mktemp("prop") / "f.pt"
        with fastsave.slot_ranges([(hb.data_ptr, hb.capacity, hb)]):
            assert fastsave.save(obj, path) == "parallel"
        with zipfile.ZipFile(path) as z:
            assert z.testzip() is None
        loaded = torch.load(path, mmap=True)
        got = loaded["slot"]
        assert all(torch.equal(a, b) for a, b in zip(got, views))
        offs = ptzip.tensor_offsets_in_file(path, got)
        if any(sizes):
            assert offs is not None and len(offs) == len(sizes)
            raw = open(path, "rb").read()
            for o, s, v in zip(offs, sizes, views):
                assert raw[o : o + s] == bytes(v.numpy().tobytes())
        del loaded, got, views
    finally:
        hb.close()
