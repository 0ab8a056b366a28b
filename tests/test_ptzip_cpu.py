"""Payload-first checkpoint container (b200/ptzip.py): files must be ordinary PyTorch checkpoints and valid ZIPs."""
import os
import zipfile

import numpy as np
import pytest
import torch


def make_obj(hb):
    g = torch.Generator().manual_seed(1)
    specs = [((300, 700), torch.float32), ((5,), torch.int64), ((), torch.float32), ((0, 3), torch.float32), ((4097,), torch.bfloat16), ((1 << 20,), torch.uint8)]
    views, off = [], 0
    for shape, dt in specs:
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        off = (off + 511) // 512 * 512
        v = hb.segment(off, n, dt, shape)
        if v.numel():
            src = torch.randn(shape, generator=g) * 50 if dt.is_floating_point else torch.randint(0, 200, shape, generator=g)
            v.copy_(src.to(dt))
        views.append(v)
        off += n
    foreign = torch.arange(12, dtype=torch.float64)
    return {"model": {"w": views[0], "ids": views[1]}, "opt": [views[2], views[3], {"m": views[4]}], "blob": views[5],
            "foreign": foreign, "shared": views[0], "iteration": 12345, "name": "x" * 100}


def same(a, b):
    if isinstance(a, dict):
        assert list(a) == list(b)
        for k in a:
            same(a[k], b[k])
    elif isinstance(a, list):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            same(x, y)
    elif isinstance(a, torch.Tensor):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert a.numel() == 0 or torch.equal(a.contiguous().view(-1).view(torch.uint8), b.contiguous().view(-1).view(torch.uint8))
    else:
        assert a == b


@pytest.mark.parametrize("force_zip64", [False, True])
@pytest.mark.parametrize("fixed_size", [False, True])
def test_container_is_a_plain_checkpoint_and_a_valid_zip(built_library, tmp_path, force_zip64, fixed_size):
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip
    from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer

    hb = HostBuffer.create(8 << 20, pin=False, prefault_threads=2)
    obj = make_obj(hb)

    def locate(ptr, nb):
        if hb.data_ptr <= ptr and ptr + nb <= hb.data_ptr + hb.capacity:
            return hb, ptr - hb.data_ptr
        return None

    path = tmp_path / "ckpt.pt"
    size = (6 << 20) + 12345 if fixed_size else None
    layout = ptzip.save(obj, path, locate=locate, threads=4, file_size=size, force_zip64=force_zip64)
    if fixed_size:
        assert os.path.getsize(path) == size
    # 1. PyTorch reads it (copying and mmap), shared storage included
    for kw in ({}, {"mmap": True}):
        got = torch.load(path, **kw)
        same(obj, got)
        assert got["shared"].data_ptr() == got["model"]["w"].data_ptr()
    # 2. it is a valid ZIP: every record stored, CRCs correct, payload where the plan said
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None
        infos = {i.filename: i for i in z.infolist()}
        assert all(i.compress_type == zipfile.ZIP_STORED for i in infos.values())
        for rec in layout.records:
            assert infos[rec.name].file_size == rec.size and infos[rec.name].header_offset == rec.header_off
            assert rec.data_off % 64 == 0
        assert "ckpt/data.pkl" in infos and "ckpt/version" in infos
    # 3. the payload region depends only on the storage sizes: a different skeleton leaves it untouched
    obj2 = dict(obj, iteration=7, name="y", extra=list(range(1000)))
    small2, stor2 = ptzip.describe(obj2)
    lay2 = ptzip.plan_payload("ckpt", [nb for _, nb in stor2], force_zip64=force_zip64)
    assert [(r.header_off, r.data_off, r.size) for r in lay2.records] == [(r.header_off, r.data_off, r.size) for r in layout.records]
    del got, obj, obj2
    hb.close()


def test_payload_can_be_written_before_the_container(built_library, tmp_path):
    """Zero-copy order of events: the payload lands first (as a DMA would), headers / pickle / directory are added after."""
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    a, b = torch.arange(1000, dtype=torch.float32), torch.arange(7, dtype=torch.int16)
    obj = {"a": a, "nested": {"b": b}, "step": 3}
    small, storages = ptzip.describe(obj)
    layout = ptzip.plan_payload("late", [nb for _, nb in storages])
    path = tmp_path / "late.pt"
    total = layout.end + ptzip.tail_size("late", small, len(storages)) + 4096
    fd = os.open(path, os.O_CREAT | os.O_RDWR, 0o644)
    os.ftruncate(fd, total)
    for rec, t in zip(layout.records, (a, b)):  # "the drain"
        os.pwrite(fd, t.numpy().tobytes(), rec.data_off)
    end = ptzip.write_container(fd, layout, small, file_size=total)
    os.close(fd)
    assert end == total
    got = torch.load(path)
    assert torch.equal(got["a"], a) and torch.equal(got["nested"]["b"], b) and got["step"] == 3
    with pytest.raises(ValueError):
        fd = os.open(tmp_path / "small.pt", os.O_CREAT | os.O_RDWR, 0o644)
        try:
            ptzip.write_container(fd, layout, small, file_size=layout.end + 10)
        finally:
            os.close(fd)


def test_patch_record_crcs_also_in_zip64_containers(tmp_path):
    """ptzip.patch_record_crcs on a container PyTorch laid out with skip_data: data descriptors and central directory take
    the checksums (small file: zipfile.testzip passes afterwards; > 4 GiB sparse file: the zip64 directory is found and patched)."""
    import os
    import zipfile
    import zlib

    import torch

    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    sd = {"a": torch.arange(1000, dtype=torch.float32), "b": {"c": torch.ones(33, dtype=torch.int64)}, "z": torch.empty(0)}
    path = tmp_path / "small.pt"
    with torch.serialization.skip_data():
        torch.save(sd, path)
    reader = torch._C.PyTorchFileReader(str(path))
    recs, crcs = [], []
    fd = os.open(path, os.O_RDWR)
    try:
        for name, t in (("data/0", sd["a"]), ("data/1", sd["b"]["c"])):
            raw = t.numpy().tobytes()
            off = reader.get_record_offset(name)
            os.pwrite(fd, raw, off)
            recs.append((name, off, len(raw)))
            crcs.append(zlib.crc32(raw))
        assert ptzip.patch_record_crcs(fd, recs, crcs) == 2
    finally:
        os.close(fd)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None
    got = torch.load(path)
    assert torch.equal(got["a"], sd["a"]) and torch.equal(got["b"]["c"], sd["b"]["c"])

    big = {"big": torch.empty(4 * 1024**3 + 4096, dtype=torch.uint8), "tail": torch.empty(1 << 20, dtype=torch.uint8)}  # never touched
    path = tmp_path / "big.pt"
    with torch.serialization.skip_data():
        torch.save(big, path)
    reader = torch._C.PyTorchFileReader(str(path))
    recs = [(n, reader.get_record_offset(n), big[k].numel()) for n, k in (("data/0", "big"), ("data/1", "tail"))]
    assert recs[1][1] > 4 * 1024**3
    fd = os.open(path, os.O_RDWR)
    try:
        assert ptzip.patch_record_crcs(fd, recs, [0x11223344, 0x55667788]) == 2
    finally:
        os.close(fd)
    with zipfile.ZipFile(path) as z:
        by_name = {zi.filename.split("/", 1)[1]: zi for zi in z.infolist()}
        assert by_name["data/0"].CRC == 0x11223344 and by_name["data/1"].CRC == 0x55667788
        assert by_name["data/1"].header_offset > 4 * 1024**3


def test_mapping_base_is_read_from_proc_maps(tmp_path):
    """tensor_offsets_in_file takes the address a checkpoint is mapped at from /proc/self/maps (exact), also through a second
    name of the same file (a hard link, as the zero-copy publish creates) and for a single tensor behind another storage --
    the case where inferring the base from the records alone is ambiguous."""
    import os

    import torch

    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    obj = {"common": {"rng": torch.arange(40, dtype=torch.int64)}, "payload": [torch.arange(80, dtype=torch.int32)]}
    path = tmp_path / "one.pt"
    torch.save(obj, path)
    link = tmp_path / "other_name.pt"
    os.link(path, link)
    loaded = torch.load(path, mmap=True)
    t = loaded["payload"][0]
    reader = torch._C.PyTorchFileReader(str(path))
    want = reader.get_record_offset("data/1")  # data/0 is the host tensor pickled first
    base = ptzip._mapping_base(path, t.data_ptr())
    assert base is not None and t.data_ptr() - base == want and base % 4096 == 0
    assert ptzip._mapping_base(link, t.data_ptr()) == base  # same inode under another name
    assert ptzip.tensor_offsets_in_file(path, [t]) == [want]
    assert ptzip.tensor_offsets_in_file(link, [t]) == [want]
    assert ptzip._mapping_base(path, torch.zeros(4).data_ptr()) is None  # a pointer that is not in a mapping of the file
    assert ptzip.tensor_offsets_in_file(path, [torch.zeros(4)]) is None
