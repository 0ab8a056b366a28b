"""GPU CRC-32, everything that can be checked without a GPU: the operator tables, the per-lane arithmetic of the kernel
(the product's __host__ __device__ functions compiled for the CPU by oracle/Makefile and driven lane by lane), the chunk
enumeration, and the CPU-only chaining step -- all against zlib (oracle/crc_oracle.py)."""
import ctypes as C
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT
from oracle import crc_oracle as co


@pytest.fixture(scope="module")
def lanes():
    res = subprocess.run(["make", "-C", str(ROOT / "oracle")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lib = C.CDLL(str(ROOT / "oracle" / "_build" / "libcrc_lanes.so"))
    lib.crc_lanes_chunk_value.restype = C.c_int
    for fn in (lib.crc_lanes_chunk_value, lib.crc_lanes_chunk_value_private):
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    return lib


def operator(which):
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi

    out = (C.c_uint32 * 1024)()
    _cabi.check(_cabi.lib().nvrx_crc_operator(which, out), "nvrx_crc_operator")
    return np.frombuffer(out, dtype=np.uint32).copy()


def test_operator_tables_are_feed_n_zero_bytes(built_library):
    for which, n in ((4, 4), (16, 16), (512, 512), (0, co.ROW_BYTES * co.CHUNK_ROWS)):
        assert operator(which).tolist() == co.operator_table(n), which


def test_lane_arithmetic_gives_the_chunk_value(built_library, lanes):
    z512, z4, z16 = operator(512), operator(4), operator(16)
    rng = np.random.default_rng(3)
    for rows in (1, 2, 3, 4, 5, 7, 8, 31, 64, 127, 128):
        data = rng.integers(0, 256, rows * 512, dtype=np.uint8)
        if rows == 2:
            data[:] = 0xFF
        if rows == 3:
            data[:] = 0
        out = C.c_uint32()
        rc = lanes.crc_lanes_chunk_value(data.ctypes.data, rows, z512.ctypes.data, z4.ctypes.data, z16.ctypes.data, C.byref(out))
        assert rc == 0 and out.value == co.chunk_value(data.tobytes()), rows
        # the default kernel's table layout: one copy of Z(512) per lane, entry-major
        z512x32 = np.repeat(z512, 32)
        out2 = C.c_uint32()
        rc = lanes.crc_lanes_chunk_value_private(data.ctypes.data, rows, z512x32.ctypes.data, z4.ctypes.data, z16.ctypes.data, C.byref(out2))
        assert rc == 0 and out2.value == out.value, rows


def finish(offsets, nbytes, values, payload):
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi

    n = len(offsets)
    offs, sizes = (C.c_uint64 * max(n, 1))(*offsets), (C.c_uint64 * max(n, 1))(*nbytes)
    vals = (C.c_uint32 * max(len(values), 1))(*values)
    out = (C.c_uint32 * max(n, 1))()
    rc = _cabi.lib().nvrx_crc_finish(n, offs, sizes, vals, len(values), payload.ctypes.data, out)
    return rc, list(out)[:n]


def test_finish_chains_values_and_tails_into_zlib_crcs(built_library):
    rng = np.random.default_rng(5)
    sizes = [0, 1, 4, 511, 512, 513, 1024 + 17, 65536, 65536 + 512, 3 * 65536 + 5 * 512 + 100, 700, 2_000_003]
    offsets, cur = [], 0
    for i, nb in enumerate(sizes):
        cur = -(-cur // 512) * 512 + (8 if i == 10 else 0)  # extent 10 starts unaligned: the host sums all of it
        offsets.append(cur)
        cur += nb
    payload = rng.integers(0, 256, cur + 64, dtype=np.uint8)
    chunks = co.chunks_of(offsets, sizes)
    assert all(ext != 10 for _, _, ext in chunks) and sum(r for _, r, ext in chunks if ext == 9) == 3 * 128 + 5
    values = [co.chunk_value(payload[off : off + rows * 512].tobytes()) for off, rows, _ in chunks]
    rc, crcs = finish(offsets, sizes, values, payload)
    assert rc == 0
    assert crcs == [zlib.crc32(payload[o : o + n].tobytes()) for o, n in zip(offsets, sizes)]
    # a value list that does not match the extents is refused
    assert finish(offsets, sizes, values[:-1], payload)[0] != 0
    assert finish(offsets, sizes, values + [0], payload)[0] != 0
    assert finish([], [], [], payload) == (0, [])


def test_plan_enumerates_the_same_chunks(built_library):
    from nvidia_resiliency_ext.checkpointing.b200 import _cabi

    sizes = [3 * 65536 + 100, 4, 0, 512 * 130]
    offsets = [0, 262144 + 512, 263168, 263680]
    h = C.c_void_p()
    lib = _cabi.lib()
    _cabi.check(lib.nvrx_crc_create(4, (C.c_uint64 * 4)(*offsets), (C.c_uint64 * 4)(*sizes), 0, C.byref(h)), "nvrx_crc_create")
    n = C.c_uint64()
    _cabi.check(lib.nvrx_crc_info(h, C.byref(n)), "nvrx_crc_info")
    assert n.value == len(co.chunks_of(offsets, sizes)) == 3 + 2
    lib.nvrx_crc_destroy(h)


def test_reference_written_files_carry_zlib_record_checksums(built_library):
    """Pins what "checksum parity" means: the files the REFERENCE wrote (tests/golden, made by make_golden.py) have, for every
    tensor record, the zlib crc32 of its bytes in the ZIP directory -- the value the GPU path has to reproduce."""
    import torch

    from conftest import GOLDEN
    from oracle import snapshot_oracle as orc
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    for name in ("c1_reference_async.pt", "iter_0000007_0_local.pt"):
        loaded = torch.load(GOLDEN / name, weights_only=False)
        tensors = orc.flatten_tensors(loaded.state_dict if hasattr(loaded, "state_dict") else loaded)
        crcs = ptzip.record_crcs(GOLDEN / name, len(tensors))
        assert crcs is not None, name
        for t, crc in zip(tensors, crcs):
            raw = t.contiguous().view(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
            assert crc == zlib.crc32(raw), name
