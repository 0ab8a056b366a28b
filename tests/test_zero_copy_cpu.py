"""Zero-copy persistence, host half (no GPU): a snapshot slot packed in checkpoint-container geometry is published as a
torch.load-able file by a hard link.  The GPU's part (pack at the planner's explicit offsets, drain into the slot) is stood
in for by writing the oracle's bytes at the offsets the CUDA-free planner reports."""
import os
import zipfile
import zlib

import pytest
import torch

from oracle import snapshot_oracle as orc


def _state():
    g = torch.Generator().manual_seed(7)
    return {
        "model": {"w": torch.randn(300, 700, generator=g), "ids": torch.arange(5), "empty": torch.empty(0, 3)},
        "opt": [torch.tensor(2.5), {"m": torch.randn(4097, generator=g).to(torch.bfloat16)}],
        "blob": torch.randint(0, 255, (1 << 20,), dtype=torch.uint8, generator=g),
        "iteration": 12345,
        "name": "x" * 100,
    }


def _same(a, b):
    if isinstance(a, dict):
        assert list(a) == list(b)
        for k in a:
            _same(a[k], b[k])
    elif isinstance(a, list):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
    elif isinstance(a, torch.Tensor):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert a.numel() == 0 or torch.equal(a.contiguous().view(-1).view(torch.uint8), b.contiguous().view(-1).view(torch.uint8))
    else:
        assert a == b


def _slot_with_snapshot(state, name, container=True, spare=0):
    """What the engine leaves behind after snapshot(container=...) has drained: (HostBuffer, descriptor, skeleton)."""
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer, PackedLayout, Plan, dtype_name
    from nvidia_resiliency_ext.checkpointing.b200.persist import SnapshotRef

    tensors = orc.flatten_tensors(state)
    sizes = [t.numel() * t.element_size() for t in tensors]
    offsets = ptzip.slot_offsets(sizes)[0] if container else None
    # the planner (CUDA-free) must accept the geometry and report it back unchanged
    plan = Plan([0x10000 * (i + 1) if n else 0 for i, n in enumerate(sizes)], sizes, None, device=0, staging_offsets=offsets)
    if container:
        assert list(plan.offsets) == offsets
        assert plan.staging_bytes >= ptzip.slot_offsets(sizes)[1]
    room = ptzip.slot_tail_room(len(sizes)) if container else 0
    hb = HostBuffer.create(plan.staging_bytes + room + spare, name=name, pin=False, prefault_threads=1)
    flat = hb.as_tensor(plan.staging_bytes)
    flat.fill_(0xAB)  # gaps hold junk, as a reused slot would
    for t, off, n in zip(tensors, plan.offsets, sizes):
        if n:
            flat[off : off + n] = t.contiguous().view(-1).view(torch.uint8)
    layout = PackedLayout(
        shapes=[tuple(t.shape) for t in tensors], dtypes=[dtype_name(t.dtype) for t in tensors],
        src_dtypes=[dtype_name(t.dtype) for t in tensors], offsets=list(plan.offsets), packed_nbytes=list(plan.packed_nbytes),
        total_bytes=plan.staging_bytes,
    )
    desc = {"shm_name": name, "progress_target": 0, "layout": layout, "owner_pid": -1, "owner_base": 0}
    counter = iter(range(len(tensors)))

    def hollow(x):
        if isinstance(x, dict):
            return {k: hollow(v) for k, v in x.items()}
        if isinstance(x, list):
            return [hollow(v) for v in x]
        return SnapshotRef(next(counter)) if isinstance(x, torch.Tensor) else x

    plan.close()
    return hb, desc, hollow(state)


def test_slot_file_starts_like_a_checkpoint(built_library):
    from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer

    name = f"/nvrx_zc_prefix_{os.getpid()}"
    hb = HostBuffer.create(1 << 16, name=name, pin=False, prefault_threads=1)
    try:
        with open("/dev/shm" + name, "rb") as fh:
            page = fh.read(4096)
        assert page[:4] == b"PK\x03\x04"  # what torch.load checks before anything else
        name_len, extra_len = int.from_bytes(page[26:28], "little"), int.from_bytes(page[28:30], "little")
        assert page[30 : 30 + name_len] == b".nvrx_slot" and 30 + name_len + extra_len == 4096  # next local header at 4096
        assert int.from_bytes(page[18:22], "little") == 0  # empty record
        # the progress word lives inside that header's extra field
        assert hb.progress_ptr - hb.data_ptr == -4096 + 192
    finally:
        hb.close()


@pytest.mark.parametrize("crc", ["0", "1"])
@pytest.mark.parametrize("as_file_object", [False, True])
def test_save_publishes_the_slot_by_hard_link(built_library, shm_dir, monkeypatch, crc, as_file_object):
    from nvidia_resiliency_ext.checkpointing.b200 import fastsave, ptzip
    from nvidia_resiliency_ext.checkpointing.b200.persist import save_snapshot_with_torch

    monkeypatch.setenv("NVRX_B200_ZERO_COPY", "1")
    monkeypatch.setenv("NVRX_B200_ZIP_CRC", crc)
    state = _state()
    name = f"/nvrx_zc_{os.getpid()}_{crc}{int(as_file_object)}"
    hb, desc, skeleton = _slot_with_snapshot(state, name, spare=123_456)
    slot_path = "/dev/shm" + name
    target = shm_dir / "ckpt.pt"
    try:
        assert not ptzip.slot_is_published(slot_path)
        if as_file_object:  # how LocalCheckpointManager._save calls it
            from nvidia_resiliency_ext.checkpointing.b200.engine import host_views
            from nvidia_resiliency_ext.checkpointing.b200.persist import _materialise

            obj = _materialise(skeleton, host_views(desc["layout"], hb))
            with open(target, "x+b") as fh, fastsave.slot_ranges(fastsave.ranges_for([desc], [hb])):
                assert fastsave.save(obj, fh) == "linked"
            del obj
        else:  # how TorchAsyncCheckpoint's writer calls it
            save_snapshot_with_torch(skeleton, str(target), desc)
        assert os.path.samefile(slot_path, target) and ptzip.slot_is_published(slot_path)
        assert os.stat(target).st_size == 4096 + hb.capacity
        assert not [p for p in os.listdir(shm_dir) if p != "ckpt.pt"]  # no temporary link left behind

        _same(torch.load(target, weights_only=False), state)
        _same(torch.load(target, weights_only=False, mmap=True), state)
        _same(torch.load(target, weights_only=True), state)
        # restore side: the tensors of an mmap load are located inside the file (= inside the still-pinned slot)
        mapped = orc.flatten_tensors(torch.load(target, weights_only=False, mmap=True))
        offs = ptzip.tensor_offsets_in_file(target, mapped)
        want = [4096 + o for o in desc["layout"].offsets]
        assert offs is not None and all(a == b for a, b, t in zip(offs, want, mapped) if t.numel())
        assert ptzip.tensor_offsets_in_file(target, orc.flatten_tensors(torch.load(target, weights_only=False))) is None
        with zipfile.ZipFile(target) as zf:
            names = zf.namelist()
            assert "archive/data.pkl" in names and "archive/data/0" in names and "archive/.pad" in names
            if crc == "1":  # read() verifies the CRC of a record (torch.load never does, hence the default of none)
                for n in names:
                    if n != "archive/.pad":
                        zf.read(n)
                assert zf.read("archive/data/1") == state["model"]["ids"].numpy().tobytes()

        # the slot is idle again as soon as the checkpoint is deleted
        os.unlink(target)
        assert not ptzip.slot_is_published(slot_path)
    finally:
        hb.close()


def test_falls_back_to_the_copying_writer(built_library, tmp_path, shm_dir, monkeypatch):
    """Other file system (no hard link possible), dense geometry, or a tail that does not fit: the file is still written,
    by the parallel copy path."""
    from nvidia_resiliency_ext.checkpointing.b200 import fastsave, ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import host_views
    from nvidia_resiliency_ext.checkpointing.b200.persist import _materialise

    monkeypatch.setenv("NVRX_B200_ZERO_COPY", "1")
    state = _state()
    cases = []
    if os.stat(tmp_path).st_dev != os.stat("/dev/shm").st_dev:
        cases.append(("other-fs", True, tmp_path / "a.pt", None))
    cases.append(("dense-geometry", False, shm_dir / "b.pt", None))
    cases.append(("no-tail-room", True, shm_dir / "c.pt", 64))
    for label, container, target, shrink_tail in cases:
        name = f"/nvrx_zc_fb_{os.getpid()}_{label}"
        if shrink_tail is not None:
            monkeypatch.setattr(ptzip, "slot_tail_room", lambda n: shrink_tail)
        hb, desc, skeleton = _slot_with_snapshot(state, name, container=container)
        try:
            obj = _materialise(skeleton, host_views(desc["layout"], hb))
            with fastsave.slot_ranges(fastsave.ranges_for([desc], [hb])):
                assert fastsave.save(obj, str(target)) == "parallel", label
            assert not ptzip.slot_is_published("/dev/shm" + name), label
            _same(torch.load(target, weights_only=False), state)
        finally:
            hb.close()


def test_slot_choice_skips_published_checkpoints(built_library, shm_dir):
    from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer, _Slot, choose_slot

    slots = []
    try:
        for i in range(3):
            s = _Slot(i)
            s.buf = HostBuffer.create(1 << 16, name=f"/nvrx_zc_pool_{os.getpid()}_{i}", pin=False, prefault_threads=1)
            slots.append(s)
        link = lambda i: os.link("/dev/shm" + slots[i].buf.name, shm_dir / f"iter{i}.pt")  # noqa: E731
        assert choose_slot(slots, 100, 4) == (slots[0], False)
        link(0)
        assert slots[0].published() and choose_slot(slots, 100, 4) == (slots[1], False)
        slots[1].busy = True
        assert choose_slot(slots, 100, 4) == (slots[2], False)
        assert choose_slot(slots, 1 << 20, 4) == (slots[2], False)  # too small: the caller re-creates its buffer
        link(2)
        assert choose_slot(slots, 100, 4) == (None, False)  # room to grow the pool
        assert choose_slot(slots, 100, 3) == (slots[0], True)  # pool at its bound: give a kept checkpoint up
        os.unlink(shm_dir / "iter0.pt")
        assert choose_slot(slots, 100, 3) == (slots[0], False)  # deleting the checkpoint frees its slot
        slots[0].busy = slots[2].busy = True
        assert choose_slot(slots, 100, 3) == (None, False)  # everything in flight: the engine raises
        # headroom rule of snapshot(): publish only while another unpublished slot (or room to grow) remains
        from nvidia_resiliency_ext.checkpointing.b200.engine import spare_slots

        for s in slots:
            s.busy = False
        assert spare_slots(slots, 4) == 2 + 1  # slot 2 is published (link above), 0 and 1 are free, one more may be created
        link(0)
        assert spare_slots(slots, 4) == 1 + 1 and spare_slots(slots, 3) == 1  # at 1 the engine stops publishing
        os.unlink(shm_dir / "iter0.pt")
        os.unlink(shm_dir / "iter2.pt")
        assert spare_slots(slots, 3) == 3
    finally:
        for s in slots:
            s.buf.close()


def test_published_records_carry_the_gpus_checksums(built_library, shm_dir, monkeypatch):
    """Writer half of the GPU CRC path: partial values (here: from the oracle) and the ready word sit behind the payload; the
    published file must be a ZIP whose every record passes its CRC check, with no CPU checksum pass configured."""
    import ctypes as C
    import threading
    import time

    import numpy as np

    from oracle import crc_oracle as co
    from nvidia_resiliency_ext.checkpointing.b200.persist import save_snapshot_with_torch

    monkeypatch.setenv("NVRX_B200_ZERO_COPY", "1")
    monkeypatch.setenv("NVRX_B200_ZIP_CRC", "0")
    state = _state()
    name = f"/nvrx_zc_gpucrc_{os.getpid()}"
    hb, desc, skeleton = _slot_with_snapshot(state, name, spare=1 << 16)
    try:
        lay = desc["layout"]
        chunks = co.chunks_of(lay.offsets, lay.packed_nbytes)
        payload = hb.as_tensor(hb.capacity).numpy()
        values = np.array([co.chunk_value(payload[o : o + r * 512].tobytes()) for o, r, _ in chunks], dtype=np.uint32)
        crc_off = -(-lay.total_bytes // 64) * 64
        ready_off = crc_off + -(-4 * len(values) // 8) * 8
        desc["crc"] = {"offset": crc_off, "n_values": len(values), "ready_offset": ready_off, "ready_value": 77}

        def gpu():  # values and the ready word arrive a little after the payload, as from the checksum stream
            time.sleep(0.2)
            payload[crc_off : crc_off + 4 * len(values)] = values.view(np.uint8)
            C.c_uint64.from_address(hb.data_ptr + ready_off).value = 77

        t = threading.Thread(target=gpu)
        t.start()
        target = shm_dir / "ckpt.pt"
        save_snapshot_with_torch(skeleton, str(target), desc)
        t.join()
        assert os.path.samefile("/dev/shm" + name, target)
        # without the hard link (zero-copy off, or another file system) the same checksums go into a copied container
        monkeypatch.setenv("NVRX_B200_ZERO_COPY", "0")
        copied = shm_dir / "copied.pt"
        save_snapshot_with_torch(skeleton, str(copied), desc)
        assert not os.path.samefile("/dev/shm" + name, copied) and os.stat(copied).st_nlink == 1
        for path in (target, copied):
            _check_archive(path, state)
    finally:
        hb.close()


def _check_archive(target, state):
    if True:
        _same(torch.load(target, weights_only=False), state)
        _same(torch.load(target, weights_only=False, mmap=True), state)
        archive = "archive" if os.stat(target).st_nlink > 1 else os.path.splitext(os.path.basename(target))[0]
        with zipfile.ZipFile(target) as zf:
            for n in zf.namelist():
                if not n.endswith("/.pad"):
                    zf.read(n)  # raises BadZipFile on a wrong CRC
            tensors = orc.flatten_tensors(state)
            for i, t_ in enumerate(tensors):
                assert zf.getinfo(f"{archive}/data/{i}").CRC == zlib.crc32(t_.contiguous().view(-1).view(torch.uint8).numpy().tobytes() if t_.numel() else b"")


def test_record_crcs_reads_the_directory(built_library, tmp_path):
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    state = _state()
    tensors = orc.flatten_tensors(state)
    path = tmp_path / "plain.pt"
    torch.save(state, path)  # PyTorch's own writer fills the CRC fields in
    want = [zlib.crc32(t.contiguous().view(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b"") for t in tensors]
    assert ptzip.record_crcs(path, len(tensors)) == want
    assert ptzip.record_crcs(path, len(tensors) + 1) is None  # not one record per tensor
    (tmp_path / "junk.pt").write_bytes(b"not a zip")
    assert ptzip.record_crcs(tmp_path / "junk.pt", 1) is None and ptzip.record_crcs(tmp_path / "missing.pt", 1) is None


def test_slots_of_dead_owners_are_reaped(tmp_path):
    """A killed trainer cannot unlink its slots; the next engine start removes them -- but only names of THIS pid namespace
    whose pid is gone, and never a checkpoint that was published from such a slot."""
    import subprocess
    import sys

    from nvidia_resiliency_ext.checkpointing.b200.engine import _pid_namespace, reap_stale_slots

    dead = subprocess.Popen([sys.executable, "-c", "pass"])
    dead.wait()
    ns = _pid_namespace()
    names = {
        "dead": f"nvrx_b200_{ns}_{dead.pid}_0123abcd_s0_g1",
        "dead_published": f"nvrx_b200_{ns}_{dead.pid}_0123abcd_s1_g2",
        "alive": f"nvrx_b200_{ns}_{os.getpid()}_0123abcd_s0_g1",
        "other_namespace": f"nvrx_b200_{ns + 1}_{dead.pid}_0123abcd_s0_g1",
        "not_a_slot": f"nvrx_b200_test_{dead.pid}_whatever",
    }
    for n in names.values():
        (tmp_path / n).write_bytes(b"x" * 10)
    os.link(tmp_path / names["dead_published"], tmp_path / "iter_0000001_0_local.pt")
    removed = reap_stale_slots(str(tmp_path))
    assert sorted(removed) == sorted([names["dead"], names["dead_published"]])
    left = set(os.listdir(tmp_path))
    assert left == {names["alive"], names["other_namespace"], names["not_a_slot"], "iter_0000001_0_local.pt"}
    assert (tmp_path / "iter_0000001_0_local.pt").read_bytes() == b"x" * 10 and os.stat(tmp_path / "iter_0000001_0_local.pt").st_nlink == 1
