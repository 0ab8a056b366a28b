"""Pins the oracle (oracle/snapshot_oracle.py) and the product's host logic to the REFERENCE's own outputs
(tests/golden/*, produced by tests/golden/make_golden.py from /root/reference).  CPU only."""
import json
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import snapshot_oracle as orc


def _eq(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a.view(torch.uint8) if a.numel() else a, b.view(torch.uint8) if b.numel() else b)


def bit_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.numel() == 0:
        return True
    return torch.equal(a.contiguous().view(-1).view(torch.uint8), b.contiguous().view(-1).view(torch.uint8))


def assert_same_tree(x, y, path=()):
    assert type(x) is type(y) or (isinstance(x, (int, float)) and isinstance(y, (int, float))), path
    if isinstance(x, dict):
        assert list(x.keys()) == list(y.keys()), path
        for k in x:
            assert_same_tree(x[k], y[k], path + (k,))
    elif isinstance(x, (list, tuple)):
        assert len(x) == len(y), path
        for i, (a, b) in enumerate(zip(x, y)):
            assert_same_tree(a, b, path + (i,))
    elif isinstance(x, torch.Tensor):
        assert bit_equal(x, y), path
    else:
        assert x == y, path


# ---- C1: the reference's async torch.save output -------------------------------------------------
def test_reference_async_file_equals_inputs_bitwise():
    inputs = torch.load(GOLDEN / "c1_inputs.pt", weights_only=False)
    ref = torch.load(GOLDEN / "c1_reference_async.pt", weights_only=False)
    assert_same_tree(inputs, ref)


def test_oracle_reproduces_reference_c1(tmp_path):
    inputs = torch.load(GOLDEN / "c1_inputs.pt", weights_only=False)
    out = tmp_path / "oracle.pt"
    orc.reference_snapshot_file(inputs, out)
    assert_same_tree(torch.load(out, weights_only=False), torch.load(GOLDEN / "c1_reference_async.pt", weights_only=False))
    # flattening order restated == order torch.save pickled the reference's dict in
    flat = orc.flatten_tensors(inputs)
    assert [tuple(t.shape) for t in flat] == [(64, 33), (33,), (17, 5), (64, 33), (), (7,), (), (3,), (251,), (0, 4)]


def test_product_queue_reproduces_reference_c1(tmp_path, dist_1rank):
    """BASELINE config C1 through the product's AsyncCallsQueue (persistent worker, CPU tensors)."""
    from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue, AsyncRequest
    from nvidia_resiliency_ext.checkpointing.utils import preload_tensors

    inputs = torch.load(GOLDEN / "c1_inputs.pt", weights_only=False)
    out = tmp_path / "product.pt"
    q = AsyncCallsQueue(persistent=True, cpu_shm_mode=True)
    idx = q.schedule_async_request(AsyncRequest(torch.save, (preload_tensors(inputs), out), [], {}))
    assert q.maybe_finalize_async_calls(blocking=True, no_dist=True) == [idx]
    q.close()
    assert_same_tree(torch.load(out, weights_only=False), torch.load(GOLDEN / "c1_reference_async.pt", weights_only=False))


# ---- the reference's own local snapshot file -------------------------------------------------------
def test_reference_local_file_loads_with_product_classes():
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict

    tasd = torch.load(GOLDEN / "iter_0000007_0_local.pt", weights_only=False)
    assert type(tasd) is BasicTensorAwareStateDict and not tasd.is_hollow
    inputs = torch.load(GOLDEN / "local_inputs.pt", weights_only=False)
    assert_same_tree(inputs, tasd.state_dict)
    got = list(tasd.tensors)
    exp = orc.flatten_tensors(inputs)
    assert len(got) == len(exp) == 5 and all(bit_equal(a, b) for a, b in zip(got, exp))


def _cpu_tasd(sd):
    from nvidia_resiliency_ext.checkpointing.local.basic_state_dict import BasicTensorAwareStateDict

    t = BasicTensorAwareStateDict.__new__(BasicTensorAwareStateDict)  # CPU tensors: skip the is_cuda assert
    t.state_dict = sd
    t._is_hollow = False
    return t


def test_product_local_manager_matches_reference_file(tmp_path, dist_1rank):
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    names = json.load(open(GOLDEN / "replication.json"))["names"]
    inputs = torch.load(GOLDEN / "local_inputs.pt", weights_only=False)
    mgr = LocalCheckpointManager(tmp_path)
    mgr.save(_cpu_tasd(inputs), 7, is_async=False)
    assert mgr.find_latest() == 7
    path = mgr._local_ckpt_path_from_id(mgr._ckpt_id(7))
    assert path.name == names["file"] == orc.local_ckpt_filename(7, 0)
    assert mgr._local_ckpt_path_from_id(mgr._ckpt_id(7), True).name == names["dirty"] == orc.local_ckpt_filename(7, 0, True)
    assert mgr._filename_from_template("\\d+", "\\d+", "\\") == names["regex"]
    assert mgr._filename_from_template("*", "*", "*") == names["glob_all"]
    assert mgr._filename_from_template(12, "*", "*") == names["glob_iter"]
    assert list(mgr._filename_to_id("iter_0000042_3_local.pt")) == names["to_id"]
    mine = torch.load(path, weights_only=False)
    ref = torch.load(GOLDEN / "iter_0000007_0_local.pt", weights_only=False)
    assert_same_tree(ref.state_dict, mine.state_dict)


# ---- clique membership / retrieve plans / coverage -------------------------------------------------
def test_parse_group_sequence_matches_reference():
    from nvidia_resiliency_ext.checkpointing.local.replication.group_utils import parse_group_sequence

    gold = json.load(open(GOLDEN / "replication.json"))["groups"]
    assert len(gold) >= 9
    for g in gold:
        exp = [tuple(x) for x in g["groups"]]
        assert orc.parse_group_sequence(g["J"], g["F"], g["W"]) == exp
        assert [tuple(x) for x in parse_group_sequence(g["J"], g["F"], g["W"])] == exp
    # the reference's docstring example (strategies.py:215-220)
    ex = orc.parse_group_sequence(8, 2, 32)
    assert ex[:3] == [(0, 8), (1, 9), (2, 10)] and ex[-1] == (23, 31) and len(ex) == 16
    with pytest.raises(AssertionError):
        parse_group_sequence(3, 2, 8)


class _StubGroup:
    def __init__(self, ranks, gathered):
        self.ranks, self._g = ranks, gathered

    def all_gather_object(self, obj):
        return self._g


def test_retrieve_plan_matches_reference():
    from nvidia_resiliency_ext.checkpointing.local.replication.strategies import (
        CliqueReplicationStrategy,
        NoReplicasAvailableError,
    )

    for sc in json.load(open(GOLDEN / "replication.json"))["plans"]:
        avail = {int(r): [tuple(i) for i in v] for r, v in sc["avail"].items()}
        wanted = [[tuple(i) for i in w] for w in sc["wanted"]]
        strat = CliqueReplicationStrategy.__new__(CliqueReplicationStrategy)
        strat.local_group = _StubGroup(sc["members"], wanted)
        if sc["entries"] is None:
            with pytest.raises(NoReplicasAvailableError):
                strat.retrieve_plan(avail, wanted[0])
            with pytest.raises(LookupError):
                orc.retrieve_plan(avail, wanted, sc["members"])
            continue
        exp = [(s, r, tuple(i)) for s, r, i in sc["entries"]]
        assert orc.retrieve_plan(avail, wanted, sc["members"]) == exp
        plan = strat.retrieve_plan(avail, wanted[0])
        assert [(e.sender, e.receiver, e.id_) for e in plan.entries] == exp


def test_find_latest_coverage_matches_reference(tmp_path, monkeypatch):
    import nvidia_resiliency_ext.checkpointing.local.ckpt_managers.base_manager as bm
    from nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager import LocalCheckpointManager

    for case in json.load(open(GOLDEN / "replication.json"))["find_latest"]:
        gathered = [[tuple(i) for i in g] for g in case["gathered"]]
        assert orc.find_latest(gathered, case["ranks"]) == case["latest"]

        class GW:
            ranks = case["ranks"]

            def all_gather_object(self, obj):
                return gathered

        monkeypatch.setattr(bm, "GroupWrapper", GW)
        mgr = LocalCheckpointManager(tmp_path, repl_strategy=object())
        mgr._rank = 0
        assert mgr.find_latest() == case["latest"]


# ---- narrowing -------------------------------------------------------------------------------------
def test_bf16_oracle_pinned_to_torch():
    z = np.load(GOLDEN / "bf16_cases.npz")
    bits, gold = z["f32_bits"], z["bf16_bits_torch_cpu"]
    nan = (bits & 0x7FFFFFFF) > 0x7F800000
    mine = orc.f32_bits_to_bf16_bits(bits)
    assert np.array_equal(mine[~nan], gold[~nan])  # RNE incl. denormals, overflow to inf, ties
    assert nan.sum() >= 5
    assert np.all(mine[nan] == orc.BF16_NAN_CUDA)
    assert np.all((gold[nan] & 0x7FFF) > 0x7F80)  # torch-CPU also yields *a* NaN there; its payload is build dependent
    # live check against this PyTorch build as well
    x = torch.from_numpy(bits.view(np.float32).copy())
    live = x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(live, gold)
    # widening is exact
    assert np.array_equal(orc.bf16_bits_to_f32_bits(gold), gold.astype(np.uint32) << 16)


# ---- packed layout specification ---------------------------------------------------------------------
def test_layout_rule_restatements_agree():
    from nvidia_resiliency_ext.checkpointing.b200.engine import expected_layout

    rng = random.Random(5)
    for _ in range(50):
        n = rng.randint(0, 40)
        nbytes = [rng.choice([0, 4, 8, 100, 512, 513, 4096, 1 << 20]) * rng.choice([1, 1, 4]) for _ in range(n)]
        narrow = [nb % 4 == 0 and rng.random() < 0.4 for nb in nbytes]
        for align in (16, 256, 512):
            assert orc.pack_layout(nbytes, narrow, align) == expected_layout(nbytes, narrow, align)
    offs, packed, total = orc.pack_layout([4, 0, 1000, 8], [False, False, True, False], 512)
    assert (offs, packed, total) == ([0, 512, 512, 1024], [4, 0, 500, 8], 1536)


def test_pack_scatter_oracle_roundtrip():
    g = torch.Generator().manual_seed(9)
    ts = [torch.randn(5, 7, generator=g), torch.randint(0, 99, (13,), generator=g, dtype=torch.int64), torch.empty(0),
          torch.tensor(3.5), torch.randn(9, generator=g).to(torch.bfloat16)]
    buf, offs, packed = orc.pack_oracle(ts)
    back = orc.scatter_oracle(buf, [t.shape for t in ts], [t.dtype for t in ts], offs, packed, [False] * len(ts))
    assert all(bit_equal(a, b) for a, b in zip(ts, back))
    buf, offs, packed = orc.pack_oracle(ts, narrow=True)
    back = orc.scatter_oracle(buf, [t.shape for t in ts], [torch.bfloat16 if t.dtype == torch.float32 and t.numel() else t.dtype for t in ts],
                              offs, packed, orc.narrow_mask(ts, True))
    assert bit_equal(back[0], ts[0].to(torch.bfloat16).to(torch.float32)) and bit_equal(back[1], ts[1])
    shard, bounds = orc.shard_bounds(10_000, 8)
    assert shard == 1536 and bounds[0] == (0, 1536) and bounds[-1][1] == 10_000 and bounds[6] == (9216, 10_000)


def test_utils_diff_matches_reference():
    """``checkpointing.utils.diff`` (reference utils.py:124-182, what the reference's own tests compare state dicts with):
    same only-left / only-right / mismatch reports, same exception where the reference raises."""
    from _diff_cases import cases, normal

    from nvidia_resiliency_ext.checkpointing.utils import diff

    golden = json.load(open(GOLDEN / "utils_diff.json"))
    assert set(golden) == set(cases())
    for name, (left, right) in cases().items():
        want = golden[name]
        if "raises" in want:
            with pytest.raises(Exception) as err:
                diff(left, right)
            assert type(err.value).__name__ == want["raises"], name
        else:
            assert normal(diff(left, right)) == want, name


def test_logging_helpers_and_wrap_for_async_behave_like_the_reference(caplog):
    """reference utils.py:35-82,102-120: ``debug_time`` logs "<scope path> took <s>s" (DEBUG, or WARNING once a threshold is
    given and reached), scopes nest with ".", ``debug_msg`` prefixes the scope path, ``wrap_for_async`` runs the function with
    the collector off and turns it back on."""
    import gc
    import logging

    from nvidia_resiliency_ext.checkpointing.utils import debug_msg, debug_time, wrap_for_async

    log = logging.getLogger("nvrx.test.scope")
    with caplog.at_level(logging.DEBUG, logger="nvrx.test.scope"):
        with debug_time("outer", log):
            with debug_time("inner"):
                debug_msg("hello")
            with debug_time("quick", threshold=60.0):
                pass
            with debug_time("slow", threshold=0.0):
                pass
    got = [(r.levelno, r.getMessage()) for r in caplog.records if r.name == "nvrx.test.scope"]
    assert (logging.DEBUG, "outer.inner hello") in got
    assert any(lvl == logging.DEBUG and msg.startswith("outer.inner took ") and msg.endswith("s") for lvl, msg in got)
    assert not any("outer.quick" in msg for _, msg in got)
    assert any(lvl == logging.WARNING and msg.startswith("outer.slow took ") for lvl, msg in got)
    assert any(lvl == logging.DEBUG and msg.startswith("outer took ") for lvl, msg in got)

    seen = []
    wrapped = wrap_for_async(lambda sd, path, flag=False: seen.append((gc.isenabled(), sd, path, flag)))
    assert gc.isenabled()
    assert wrapped({"a": 1}, "p", flag=True) is None
    assert seen == [(False, {"a": 1}, "p", True)] and gc.isenabled()
    gc.disable()
    try:
        wrapped({}, "q")
        assert not gc.isenabled()  # it was off before: stays off
    finally:
        gc.enable()
