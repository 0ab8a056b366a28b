"""The planner of libnvrx_snap.so (layout + tile work-list) runs without CUDA: check it against the oracle's layout rule and
against first principles -- every source byte is covered exactly once, bulk tiles are TMA-legal, sharded tile lists never
straddle a shard boundary.  (The kernels that walk these lists are tested on the GPU.)"""
import random

import pytest

from oracle import snapshot_oracle as orc


def plan_for(ptrs, nbytes, flags=None, **kw):
    from nvidia_resiliency_ext.checkpointing.b200.engine import Plan

    return Plan(ptrs, nbytes, flags, device=0, **kw)


def check_cover(plan, ptrs, nbytes, flags, tile_bytes, shard_bytes=0):
    n_bulk, tiles = plan.tiles(shard_bytes)
    covered = [[] for _ in nbytes]
    for idx, (seg, nb, off) in enumerate(tiles):
        assert 0 < nb <= tile_bytes and off + nb <= nbytes[seg]
        narrow = bool(flags and flags[seg])
        if idx < n_bulk:
            assert not narrow and nb % 16 == 0 and nb >= 1024 and (ptrs[seg] + off) % 16 == 0  # what cp.async.bulk needs
        if shard_bytes:
            scale = 2 if narrow else 1
            lo = plan.offsets[seg] + off // scale
            hi = plan.offsets[seg] + (off + nb) // scale
            assert lo // shard_bytes == (hi - 1) // shard_bytes, "tile straddles a shard boundary"
        covered[seg].append((off, off + nb))
    for seg, spans in enumerate(covered):
        spans.sort()
        cur = 0
        for lo, hi in spans:
            assert lo == cur, f"segment {seg}: gap or overlap at {lo} (expected {cur})"
            cur = hi
        assert cur == nbytes[seg]
    # both lists are sorted by staging position (the pipelined snapshot bisects them); sharded plans interleave the shards
    # (destination-major order of the fused exchange) and are only sorted inside a shard
    pos = lambda t: plan.offsets[t[0]] + (t[2] // 2 if (flags and flags[t[0]]) else t[2])  # noqa: E731
    if shard_bytes:
        for part in (tiles[:n_bulk], tiles[n_bulk:]):
            by_shard = {}
            for t in part:
                by_shard.setdefault(pos(t) // shard_bytes, []).append(pos(t))
            assert all(v == sorted(v) for v in by_shard.values())
        return
    assert [pos(t) for t in tiles[:n_bulk]] == sorted(pos(t) for t in tiles[:n_bulk])
    assert [pos(t) for t in tiles[n_bulk:]] == sorted(pos(t) for t in tiles[n_bulk:])
    return n_bulk, tiles


@pytest.mark.parametrize("tile_bytes", [4096, 32768, 65536])
def test_layout_and_tiles_random(built_library, tile_bytes):
    rng = random.Random(tile_bytes)
    for trial in range(30):
        n = rng.randint(0, 60)
        nbytes = [rng.choice([0, 1, 4, 15, 16, 100, 1023, 1024, 1040, 4096, 32768, 32784, 100_000, 1 << 20]) for _ in range(n)]
        ptrs = [(0x7F0000000000 + i * (4 << 20) + rng.choice([0, 0, 0, 4, 8, 2, 1])) if nb else 0 for i, nb in enumerate(nbytes)]
        flags = [1 if (nb % 4 == 0 and ptrs[i] % 4 == 0 and nb and rng.random() < 0.3) else 0 for i, nb in enumerate(nbytes)]
        align = rng.choice([16, 512, 4096])
        plan = plan_for(ptrs, nbytes, flags, align=align, tile_bytes=tile_bytes)
        offs, packed, total = orc.pack_layout(nbytes, [bool(f) for f in flags], align)
        assert list(plan.offsets) == offs and list(plan.packed_nbytes) == packed and plan.staging_bytes == total
        assert plan.algorithmic_bytes == sum(nbytes) + sum(packed)
        check_cover(plan, ptrs, nbytes, flags, tile_bytes)
        if total:
            shard = orc.shard_bounds(total, rng.randint(1, 7), 512)[0]
            check_cover(plan, ptrs, nbytes, flags, tile_bytes, shard_bytes=shard)
        # re-pointing at tensors with other alignment classes rebuilds the list consistently
        ptrs2 = [(p + 4 if p and not f else p) for p, f in zip(ptrs, flags)]
        plan.update_ptrs(ptrs2)
        check_cover(plan, ptrs2, nbytes, flags, tile_bytes)
        plan.close()


def test_c2_shape_tile_counts(built_library):
    """The BASELINE C2 state: 1164 large tensors -> bulk tiles only, 291 4-byte steps -> 291 ragged tiles."""
    from bench import llama3_8b_shard_shapes

    sizes = []
    for _, shp in llama3_8b_shard_shapes():
        n = 1
        for d in shp:
            n *= d
        sizes.append(n * 4)
    nbytes, ptrs, cur = [], [], 0x7E0000000000
    for nb in sizes:           # model params
        nbytes.append(nb); ptrs.append(cur); cur += (nb + 511) // 512 * 512
    for nb in sizes:           # optimizer state: main_param, exp_avg, exp_avg_sq, step
        for x in (nb, nb, nb, 4):
            nbytes.append(x); ptrs.append(cur); cur += (x + 511) // 512 * 512
    assert len(nbytes) == 1455 and sum(nbytes) == 16_060_522_496 + 291 * 4
    plan = plan_for(ptrs, nbytes)
    n_bulk, tiles = plan.tiles()
    assert len(tiles) - n_bulk == 291 and all(t[1] == 4 for t in tiles[n_bulk:])
    assert n_bulk == sum(-(-nb // 32768) for nb in nbytes if nb > 4)
    assert plan.staging_bytes == sum((nb + 511) // 512 * 512 for nb in nbytes)
    plan.close()


def test_c2_in_container_geometry(built_library):
    """Zero-copy persistence packs the same state at checkpoint-container offsets (ptzip.slot_offsets): the work-list keeps
    its shape (same bulk tiles, same ragged tiles), the buffer grows only by the header gaps, and the container's tail fits
    in the room the engine reserves."""
    from bench import llama3_8b_shard_shapes
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip

    sizes = []
    for _, shp in llama3_8b_shard_shapes():
        n = 1
        for d in shp:
            n *= d
        sizes.append(n * 4)
    nbytes, ptrs, cur = [], [], 0x7E0000000000
    for nb in sizes + [x for nb in sizes for x in (nb, nb, nb, 4)]:
        nbytes.append(nb); ptrs.append(cur); cur += (nb + 511) // 512 * 512
    offsets, span = ptzip.slot_offsets(nbytes)
    dense = plan_for(ptrs, nbytes)
    at = plan_for(ptrs, nbytes, staging_offsets=offsets)
    assert list(at.offsets) == offsets and all(o % 512 == 0 for o in offsets)
    (nb_d, tiles_d), (nb_a, tiles_a) = dense.tiles(), at.tiles()
    assert nb_d == nb_a and [(t[0], t[1]) for t in tiles_d] == [(t[0], t[1]) for t in tiles_a]
    assert span <= at.staging_bytes <= dense.staging_bytes + 512 * len(nbytes) + 512
    # a generous pickle (200 bytes per tensor) still fits behind the payload
    small = [("data.pkl", b"x" * (200 * len(nbytes))), ("byteorder", b"little"), ("version", b"3\n")]
    assert ptzip.tail_size(ptzip.SLOT_ARCHIVE, small, len(nbytes)) <= ptzip.slot_tail_room(len(nbytes))
    dense.close()
    at.close()


def test_planner_rejects_bad_arguments(built_library):
    from nvidia_resiliency_ext.checkpointing.b200._cabi import SnapError

    for kw in ({"align": 24}, {"align": 8}, {"tile_bytes": 1000}, {"tile_bytes": 1 << 20}):
        with pytest.raises(SnapError):
            plan_for([0x1000], [64], None, **kw)
    with pytest.raises(SnapError):
        plan_for([0], [64])                      # null pointer with bytes
    with pytest.raises(SnapError):
        plan_for([0x1002], [64], [1])            # narrowing needs 4-byte aligned fp32
    with pytest.raises(SnapError):
        plan_for([0x1000], [66], [1])            # ... and a multiple of 4 bytes
    p = plan_for([0x1000, 0], [64, 0])
    with pytest.raises(SnapError):
        p.update_ptrs([0, 0])
    p.close()


def test_shard_interleave_spreads_every_window_over_all_shards(built_library, monkeypatch):
    """Default order of a sharded plan (fused pack + all-to-all): list position q holds a tile of shard (first + q) mod n while
    every shard still has tiles, so any window of CTAs stores to all destination GPUs at once; same tiles as the plain order."""
    monkeypatch.delenv("NVRX_B200_SHARD_INTERLEAVE", raising=False)
    nbytes = [3 << 20, 100, 5 << 20, 4, 1 << 20, 7 << 20]
    ptrs = [0x7F0000000000 + i * (16 << 20) for i in range(len(nbytes))]
    plan = plan_for(ptrs, nbytes, tile_bytes=32768)
    total = plan.staging_bytes
    n = 7
    shard = orc.shard_bounds(total, n, 512)[0]
    pos = lambda t: plan.offsets[t[0]] + t[2]  # noqa: E731
    monkeypatch.setenv("NVRX_B200_SHARD_INTERLEAVE", "0")
    nb0, base = plan.tiles(shard)
    monkeypatch.delenv("NVRX_B200_SHARD_INTERLEAVE")
    for first in (0, 3, 6):
        plan.set_shard_rotation(first)
        nb, tiles = plan.tiles(shard)
        assert nb == nb0 and sorted(tiles) == sorted(base)
        shards = [pos(t) // shard for t in tiles[:nb]]
        per_shard = min(shards.count(s) for s in range(n))
        head = shards[: per_shard * n]  # while every shard still has tiles the walk is a strict round robin
        assert head == [(first + q) % n for q in range(len(head))]
        for s in range(n):  # inside a shard the tiles keep their staging order
            mine = [pos(t) for t in tiles[:nb] if pos(t) // shard == s]
            assert mine == sorted(mine)
    plan.set_shard_rotation(0)
    plan.close()


def test_shard_rotation_only_reorders(built_library, monkeypatch):
    """nvrx_plan_set_shard_rotation with NVRX_B200_SHARD_INTERLEAVE=0: same tiles, walk starts at the requested shard and wraps."""
    monkeypatch.setenv("NVRX_B200_SHARD_INTERLEAVE", "0")
    nbytes = [3 << 20, 100, 5 << 20, 4, 1 << 20]
    ptrs = [0x7F0000000000 + i * (16 << 20) for i in range(len(nbytes))]
    plan = plan_for(ptrs, nbytes, tile_bytes=32768)
    total = plan.staging_bytes
    shard = orc.shard_bounds(total, 4, 512)[0]
    nb0, base = plan.tiles(shard)
    pos = lambda t: plan.offsets[t[0]] + t[2]  # noqa: E731
    for first in (1, 2, 3):
        plan.set_shard_rotation(first)
        nb, tiles = plan.tiles(shard)
        assert nb == nb0 and sorted(tiles) == sorted(base)
        assert pos(tiles[0]) // shard == first                      # bulk list starts inside the requested shard
        shards = [pos(t) // shard for t in tiles[:nb]]
        assert shards == sorted(shards[: len(shards)], key=lambda s: (s - first) % 4)  # first, first+1, ..., wrap
    plan.set_shard_rotation(0)
    assert plan.tiles(shard)[1] == base
    plan.close()


def test_explicit_staging_offsets_follow_a_checkpoint_container(built_library):
    """nvrx_plan_create_at: the packed image can take the geometry of a payload-first container (b200/ptzip.py)."""
    from nvidia_resiliency_ext.checkpointing.b200 import ptzip
    from nvidia_resiliency_ext.checkpointing.b200._cabi import SnapError

    nbytes = [4 << 20, 4, 100_000, 0, 1 << 20]
    ptrs = [(0x7F0000000000 + i * (16 << 20)) if nb else 0 for i, nb in enumerate(nbytes)]
    lay = ptzip.plan_payload("ckpt", nbytes)
    offs = [r.data_off for r in lay.records]
    plan = plan_for(ptrs, nbytes, staging_offsets=offs)
    assert list(plan.offsets) == offs and list(plan.packed_nbytes) == nbytes
    assert plan.staging_bytes >= lay.end and plan.staging_bytes % 512 == 0
    check_cover(plan, ptrs, nbytes, None, 32768)
    plan.close()
    with pytest.raises(SnapError):
        plan_for(ptrs, nbytes, staging_offsets=[o + 8 for o in offs])      # not 16-byte aligned
    with pytest.raises(SnapError):
        plan_for(ptrs, nbytes, staging_offsets=[0, 64, 128, 256, 512])     # overlapping


def test_exchange_geometry_is_computed_once_per_structure(built_library):
    """b200/exchange._geometry: per-member layouts inside the exchange buffer (the oracle's layout rule, shifted by the member's
    slice) and their union; the result is reused while the gathered placeholder lists are the same object."""
    import torch

    from nvidia_resiliency_ext.checkpointing.b200.exchange import _geometry
    from nvidia_resiliency_ext.checkpointing.local.replication.torch_device_utils import TensorPlaceholder

    class Group:  # stands in for GroupWrapper: only its attribute dict is used
        pass

    members = [
        [torch.empty(5, 7), torch.empty(3, dtype=torch.int64), torch.empty(0), torch.empty((), dtype=torch.bfloat16)],
        [torch.empty(1000, 3), torch.empty(2, dtype=torch.uint8), torch.empty(4, 4), torch.empty(1)],
    ]
    placeholders = [[TensorPlaceholder(t) for t in tensors] for tensors in members]
    assert [tp.nbytes for tp in placeholders[1]] == [12000, 2, 64, 4]
    group = Group()
    geo = _geometry(group, placeholders, 512, 2)
    slot_bytes = geo["slot_bytes"]
    assert slot_bytes % 512 == 0
    for r, tensors in enumerate(members):
        offs, packed, total = orc.pack_layout([t.numel() * t.element_size() for t in tensors], [False] * len(tensors), 512)
        assert geo["layouts"][r] == (offs, packed, total) and total <= slot_bytes
        lay = geo["dev_lists"][r]
        assert lay.offsets == [r * slot_bytes + o for o in offs] and lay.packed_nbytes == packed
        assert lay.shapes == [tuple(t.shape) for t in tensors] and lay.total_bytes == 2 * slot_bytes
    union = geo["union"]
    assert union.offsets == geo["dev_lists"][0].offsets + geo["dev_lists"][1].offsets and len(union.shapes) == 8
    assert _geometry(group, placeholders, 512, 2) is geo            # same structure object: cached
    assert _geometry(group, [list(p) for p in placeholders], 512, 2) is not geo  # a new gather result: recomputed


def test_exchange_geometry_in_container_mode(built_library):
    """Replicated zero-copy: every member's slice of the exchange buffer carries the container geometry of that member's
    tensors, so that it can be drained into its own slot and published as that member's file."""
    import torch

    from nvidia_resiliency_ext.checkpointing.b200 import ptzip
    from nvidia_resiliency_ext.checkpointing.b200.exchange import _geometry
    from nvidia_resiliency_ext.checkpointing.local.replication.torch_device_utils import TensorPlaceholder

    class Group:
        pass

    members = [[torch.empty(5, 7), torch.empty(0), torch.empty(3, dtype=torch.int64)], [torch.empty(1000, 3), torch.empty(1)]]
    placeholders = [[TensorPlaceholder(t) for t in tensors] for tensors in members]
    group = Group()
    dense = _geometry(group, placeholders, 512, 2)
    geo = _geometry(group, placeholders, 512, 2, container=True)
    assert geo is not dense and geo["container"] and _geometry(group, placeholders, 512, 2, container=True) is geo
    for r, tensors in enumerate(members):
        sizes = [t.numel() * t.element_size() for t in tensors]
        offs, span = ptzip.slot_offsets(sizes)
        assert geo["layouts"][r][0] == offs and geo["layouts"][r][2] >= span
        assert geo["own_lists"][r].offsets == offs and geo["own_lists"][r].total_bytes == geo["layouts"][r][2]
        assert geo["dev_lists"][r].offsets == [r * geo["slot_bytes"] + o for o in offs]
        # the planner of member r produces exactly these offsets for its own tensors
        plan = plan_for([0x1000000 * (i + 1) if n else 0 for i, n in enumerate(sizes)], sizes, staging_offsets=offs)
        assert list(plan.offsets) == offs and plan.staging_bytes == geo["layouts"][r][2]
        plan.close()
    assert geo["slot_bytes"] >= max(l[2] for l in geo["layouts"]) and geo["slot_bytes"] % 512 == 0
