"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's checkpoint-snapshot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the *checker* or the *timed baseline*; nothing under
``nvidia-resiliency-ext_b200/`` imports it (tests/test_no_oracle_in_product.py enforces that).

Reference = NVIDIA/nvidia-resiliency-ext @ 02ee37aa (``/root/reference``); paths below are relative to
``src/nvidia_resiliency_ext/checkpointing``.  The reference is pure Python on top of PyTorch, so the
restatement is numpy / torch-on-CPU; every function cites the lines it follows.

PINNING (how this oracle is tied to the real reference; see tests/golden/make_golden.py):
  * snapshot bits           -- the reference itself was run in the build container (CPU path, config C1:
                               ``AsyncCallsQueue(persistent=True, cpu_shm_mode=True)`` + ``torch.save``) and
                               through ``LocalCheckpointManager.save``; its output files are committed under
                               ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this module and the
                               product's loaders against them bit for bit.
  * clique membership, retrieve plans, coverage rule, file names
                            -- golden JSON produced by calling the reference's own functions.
  * fp32->bf16 narrowing    -- DOES NOT EXIST in the reference ("parity unpinned" there); defined as
                               ``x.to(torch.bfloat16)`` and pinned against PyTorch (CPU for finite values and
                               infinities; the NaN *payload* follows the CUDA ``cvt.rn.bf16.f32`` rule 0x7FFF,
                               which is what PyTorch produces on CUDA tensors -- PyTorch-CPU is not even
                               self-consistent there: 0xFFFF from its AVX-512 path, 0x7FC0 from the scalar one).
  * packed staging layout   -- new in this engine (the reference has no packed buffer); the restatement here
                               is the specification in include/nvrx_snap.h, checked against the C planner.
"""

from __future__ import annotations

import random
import time
from collections import defaultdict
from typing import Any, Dict, Iterable, List, Mapping, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------------
# flattening order and the reference snapshot
# ----------------------------------------------------------------------------------------------------


def nested_values(x):
    """DFS over dict values / list items -- the flattening order (local/basic_state_dict.py:34-41)."""
    it = x.values() if isinstance(x, dict) else x
    for v in it:
        if isinstance(v, (dict, list)):
            yield from nested_values(v)
        else:
            yield v


def flatten_tensors(state_dict) -> List[torch.Tensor]:
    """Tensors of a nested state dict in TensorAwareStateDict order (local/basic_state_dict.py:112-120)."""
    return [v for v in nested_values(state_dict) if isinstance(v, torch.Tensor)]


def map_outplace(f, x):
    """Out-of-place map over dicts/lists (utils.py:185-192)."""
    if isinstance(x, dict):
        return {k: map_outplace(f, v) for k, v in x.items()}
    if isinstance(x, list):
        return [map_outplace(f, v) for v in x]
    return f(x)


def reference_preload(state_dict, non_blocking: bool = True):
    """What the reference hands to its writer: per-tensor ``detach().to("cpu", non_blocking)`` (utils.py:85-99).
    A bit copy of every tensor; non-tensors pass through."""
    return map_outplace(
        lambda v: v.detach().to("cpu", non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v, state_dict
    )


def reference_snapshot_file(state_dict, path) -> None:
    """Synchronous equivalent of ``TorchAsyncCheckpoint.async_save`` + finalize (async_ckpt/torch_ckpt.py:43-53):
    the file the reference ends up with is ``torch.save`` of the preloaded dict."""
    sd = reference_preload(state_dict)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    torch.save(sd, path)


def tensor_bytes(t: torch.Tensor) -> np.ndarray:
    """Raw little-endian bytes of a tensor in logical (contiguous) order."""
    c = t.detach().cpu().contiguous()
    if c.numel() == 0:
        return np.zeros(0, dtype=np.uint8)
    return c.view(-1).view(torch.uint8).numpy().copy()


# ----------------------------------------------------------------------------------------------------
# fp32 -> bf16 narrowing (new option; pinned against PyTorch, see header)
# ----------------------------------------------------------------------------------------------------
BF16_NAN_CUDA = 0x7FFF  # cvt.rn.bf16.f32 canonical NaN (what torch .to(bfloat16) yields on CUDA)
BF16_NAN_TORCH_CPU_SCALAR = 0x7FC0  # c10::BFloat16 round_to_nearest_even on the host (vectorised path: 0xFFFF)


def f32_bits_to_bf16_bits(u32: np.ndarray, nan_bits: int = BF16_NAN_CUDA) -> np.ndarray:
    """Round-to-nearest-even truncation of fp32 bit patterns to bf16 bit patterns."""
    u = u32.astype(np.uint64)
    is_nan = (u & 0x7FFFFFFF) > 0x7F800000
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return np.where(is_nan, np.uint64(nan_bits), rounded).astype(np.uint16)


def bf16_bits_to_f32_bits(u16: np.ndarray) -> np.ndarray:
    """Exact widening."""
    return u16.astype(np.uint32) << 16


# ----------------------------------------------------------------------------------------------------
# packed staging layout (specification: include/nvrx_snap.h)
# ----------------------------------------------------------------------------------------------------
DEFAULT_ALIGN = 512


def pack_layout(nbytes: Sequence[int], narrow: Sequence[bool], align: int = DEFAULT_ALIGN):
    """(offsets, packed_nbytes, total): segment i starts at round_up(end of i-1, align); a narrowed fp32 segment
    occupies half its source bytes; total is rounded up to ``align``."""
    offs, packed, cur = [], [], 0
    for nb, nr in zip(nbytes, narrow):
        cur = -(-cur // align) * align
        offs.append(cur)
        pk = nb // 2 if nr else nb
        packed.append(pk)
        cur += pk
    return offs, packed, -(-cur // align) * align


def narrow_mask(tensors: Sequence[torch.Tensor], narrow: bool) -> List[bool]:
    return [bool(narrow) and t.dtype == torch.float32 and t.numel() > 0 for t in tensors]


def pack_oracle(tensors: Sequence[torch.Tensor], narrow: bool = False, align: int = DEFAULT_ALIGN,
                nan_bits: int = BF16_NAN_CUDA) -> Tuple[np.ndarray, List[int], List[int]]:
    """Expected content of the staging buffer after packing ``tensors`` (gap bytes are zero: staging is
    zero-filled at allocation and the kernels never write gaps)."""
    mask = narrow_mask(tensors, narrow)
    nbytes = [t.numel() * t.element_size() for t in tensors]
    offs, packed, total = pack_layout(nbytes, mask, align)
    buf = np.zeros(total, dtype=np.uint8)
    for t, off, pk, nr in zip(tensors, offs, packed, mask):
        if pk == 0:
            continue
        raw = tensor_bytes(t)
        if nr:
            raw = f32_bits_to_bf16_bits(raw.view(np.uint32), nan_bits).view(np.uint8)
        buf[off : off + pk] = raw
    return buf, offs, packed


def scatter_oracle(buf: np.ndarray, shapes, dtypes, offs, packed, widen: Sequence[bool]) -> List[torch.Tensor]:
    """Inverse of ``pack_oracle``: typed tensors cut out of a packed buffer; ``widen[i]`` turns bf16 payload into
    fp32 exactly."""
    out = []
    for shape, dt, off, pk, wd in zip(shapes, dtypes, offs, packed, widen):
        raw = buf[off : off + pk]
        if wd:
            raw = bf16_bits_to_f32_bits(raw.view(np.uint16)).view(np.uint8)
            dt = torch.float32
        if raw.size == 0:
            out.append(torch.empty(shape, dtype=dt))
        else:
            out.append(torch.from_numpy(raw.copy()).view(dt).view(shape))
    return out


def shard_bounds(total_bytes: int, n_peers: int, align: int = DEFAULT_ALIGN) -> Tuple[int, List[Tuple[int, int]]]:
    """Sharded replica layout of the packed range: ``n_peers`` equal shards of ``shard_bytes`` (multiple of
    ``align``) covering ``[0, total_bytes)``; shard j is stored by clique member j."""
    shard = -(-total_bytes // n_peers)
    shard = max(align, -(-shard // align) * align)
    return shard, [(j * shard, min(total_bytes, (j + 1) * shard)) for j in range(n_peers)]


# ----------------------------------------------------------------------------------------------------
# replication logic
# ----------------------------------------------------------------------------------------------------


def parse_group_sequence(replication_jump: int, replication_factor: int, world_size: int) -> List[Tuple[int, ...]]:
    """Cliques n, n+J, ..., n+(F-1)J (local/replication/group_utils.py:120-146)."""
    assert replication_jump > 0 and replication_factor > 0
    assert world_size % (replication_jump * replication_factor) == 0
    result = []
    for modulus in range(replication_jump):
        seq = list(range(modulus, world_size, replication_jump))
        for i in range(0, len(seq), replication_factor):
            result.append(tuple(seq[i : i + replication_factor]))
    result.sort()
    return result


def clique_replicate_result(ids_by_rank: Mapping[int, Any], clique: Sequence[int]) -> List[Any]:
    """What ``CliqueReplicationStrategy.replicate`` returns on every member of ``clique``: the members' state dicts
    / ids in *group-rank order*, own one included (strategies.py:113-135, group_utils.py:359-375)."""
    return [ids_by_rank[r] for r in clique]


def retrieve_plan(globally_available_ids: Mapping[int, Iterable], wanted_by_rank: Sequence[Sequence], members: Sequence[int]):
    """List of (sender, receiver, id) (strategies.py:143-179): receivers in group order; a holder serves itself,
    otherwise ``random.Random(0).choice(sorted(holders))`` with ONE generator shared by the whole plan."""
    rng = random.Random(0)
    entries = []
    for receiver, wanted in zip(members, wanted_by_rank):
        for wid in wanted:
            holders = {r for r in members if wid in globally_available_ids[r]}
            if not holders:
                raise LookupError(f"No replicated copies for id={wid} found!")
            sender = receiver if receiver in holders else rng.choice(sorted(holders))
            entries.append((sender, receiver, wid))
    return entries


def find_latest(globally_available_ids: Sequence[Iterable[Tuple[int, int, Any]]], ranks: Iterable[int]) -> int:
    """Newest iteration whose owner set equals the whole group (base_manager.py:187-202)."""
    cover = defaultdict(set)
    for ids in globally_available_ids:
        for iteration, owner, _ in ids:
            cover[iteration].add(owner)
    everyone = set(ranks)
    return max((it for it, owners in cover.items() if owners == everyone), default=-1)


def local_ckpt_filename(iteration: int, rank: int, dirty: bool = False) -> str:
    """``iter_{iteration:07d}_{rank}_local[.dirty].pt`` (local_manager.py:158-173)."""
    return f"iter_{iteration:07d}_{rank}_local{'.dirty' if dirty else ''}.pt"


# ----------------------------------------------------------------------------------------------------
# the reference's save path as a timed baseline (bench.py cpu_baseline / --impl reference)
# ----------------------------------------------------------------------------------------------------


def _child_torch_save(obj, path):
    import gc

    gc.disable()  # utils.py:102-120 / wrap_for_async
    torch.save(obj, path)


def reference_fork_save(state_dict, path, persistent: bool = False) -> Dict[str, float]:
    """Port of the reference save path with its phases timed (seconds):

    ``stall``   preload_tensors (per-tensor pinned D2H) + torch.cuda.synchronize() + fork()   -- the time
                training is blocked (async_ckpt/torch_ckpt.py:49-53, async_ckpt/core.py:345-355)
    ``total``   until the forked child finished ``torch.save`` (pickle + zip/CRC + write) and was joined
    ``d2h``     the D2H part of the stall alone
    Runs on GPU tensors when CUDA is present, otherwise on CPU tensors (reference config C1)."""
    import torch.multiprocessing as mp

    cuda = torch.cuda.is_available()
    t0 = time.perf_counter()
    sd = reference_preload(state_dict)
    if cuda:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    ctx = mp.get_context("fork")
    proc = ctx.Process(target=_child_torch_save, args=(sd, path))
    proc.start()
    t2 = time.perf_counter()
    proc.join()
    t3 = time.perf_counter()
    if proc.exitcode != 0:
        raise RuntimeError(f"reference writer child failed with exit code {proc.exitcode}")
    return {"d2h": t1 - t0, "stall": t2 - t0, "total": t3 - t0, "cores": 1}
