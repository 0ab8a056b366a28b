"""TEST INFRASTRUCTURE ONLY -- CPU statement of what the GPU CRC-32 path (csrc/crc_kernels.cuh, csrc/crc.cu) must produce.

The reference gets its record checksums from PyTorch's zip writer (``torch.save`` -> miniz ``mz_crc32``, called from
``async_ckpt/torch_ckpt.py:36-41`` and ``local/ckpt_managers/local_manager.py:117-122``), i.e. plain zlib CRC-32.  So the oracle
is ``zlib.crc32`` itself, plus the two intermediate quantities of the GPU algorithm expressed through it:

* ``feed_zeros(s, n)``  -- the linear map Z(n): CRC state ``s`` after n zero bytes (no init / final inversion)
* ``chunk_value(m)``    -- state reached from state 0 over the bytes ``m`` (what one warp emits for a chunk)

Both follow from ``zlib.crc32(data, start) == ~update(~start, data)``.
"""
import zlib
from typing import List, Sequence, Tuple

MASK = 0xFFFFFFFF
ROW_BYTES = 512
CHUNK_ROWS = 128


def update(state: int, data: bytes) -> int:
    """Raw register update (no pre/post inversion)."""
    return zlib.crc32(data, state ^ MASK) ^ MASK


def feed_zeros(state: int, n: int) -> int:
    return update(state, bytes(n))


def chunk_value(data: bytes) -> int:
    return update(0, data)


def chunks_of(offsets: Sequence[int], nbytes: Sequence[int]) -> List[Tuple[int, int, int]]:
    """(offset, rows, extent) in the order the GPU path enumerates chunks: 16-byte aligned extents, whole 512-byte rows,
    at most 128 rows per chunk."""
    out = []
    for i, (off, nb) in enumerate(zip(offsets, nbytes)):
        if off % 16:
            continue
        rows = nb // ROW_BYTES
        while rows:
            take = min(rows, CHUNK_ROWS)
            out.append((off, take, i))
            off += take * ROW_BYTES
            rows -= take
    return out


def operator_table(n: int) -> List[int]:
    """The 4x256 table of Z(n): entry j*256+b is Z(n)(b << 8j)."""
    return [feed_zeros(b << (8 * j), n) for j in range(4) for b in range(256)]
