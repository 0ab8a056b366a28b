"""The timed ports of the reference's flows (oracle/reference_port.py, what ``bench.py --impl reference`` runs) produce the
reference's own results: checked on CPU tensors / gloo against the committed outputs of the imported reference (tests/golden)."""
import hashlib
import json

import torch
import torch.distributed as dist

from _mp import run_ranks
from conftest import GOLDEN
from oracle import reference_port as rp
from oracle import snapshot_oracle as orc
from test_oracle_golden import assert_same_tree


def test_async_checkpoint_port_writes_the_references_c1_file(tmp_path):
    inputs = torch.load(GOLDEN / "c1_inputs.pt", weights_only=False)
    ck = rp.ReferenceAsyncCheckpoint()
    out = tmp_path / "port.pt"
    ck.async_save(inputs, out)
    assert ck.finalize(blocking=True) and ck.done()
    assert_same_tree(torch.load(out, weights_only=False), torch.load(GOLDEN / "c1_reference_async.pt", weights_only=False))
    ck.async_save(inputs, tmp_path / "skipped.pt", write=False)  # the bounded-sample steps of the bench: no child, no file
    assert ck.finalize(blocking=True) and not (tmp_path / "skipped.pt").exists()


def test_local_save_port_stores_the_tensors_the_reference_manager_stores(tmp_path):
    inputs = torch.load(GOLDEN / "local_inputs.pt", weights_only=False)
    res = rp.reference_local_save(inputs, tmp_path / "iter_0000007_0_local.pt")
    assert 0 < res["stall"] <= res["total"]
    port = torch.load(tmp_path / "iter_0000007_0_local.pt", weights_only=False)
    ref = torch.load(GOLDEN / "iter_0000007_0_local.pt", weights_only=False)  # written by the reference's LocalCheckpointManager
    assert_same_tree(port.state_dict, ref.state_dict)


def _sha(t):
    c = t.detach().cpu().contiguous()
    raw = c.view(-1).view(torch.uint8).numpy().tobytes() if c.numel() else b""
    return hashlib.sha256(str(c.dtype).encode() + str(tuple(c.shape)).encode() + raw).hexdigest()


def _w_gather(rank, world, out_dir):
    src = (GOLDEN / "make_golden.py").read_text()
    ns = {}
    exec("import torch\n" + src[src.index("def rank_tensors"):src.index("def _replicate_worker")], ns)  # noqa: S102 - fixture code of this repo
    mine = orc.flatten_tensors(ns["rank_tensors"](rank))
    rows = rp.reference_all_gather_batch(mine, None, "cpu")
    with open(f"{out_dir}/r{rank}.json", "w") as f:
        json.dump([[_sha(t) for t in row] for row in rows], f)


def test_all_gather_batch_port_matches_the_references_replicate(tmp_path):
    """replicate() of the imported reference on 2 gloo ranks (tests/golden/replicate_2rank.json) gathered these tensors."""
    run_ranks(_w_gather, 2, str(tmp_path))
    gold = json.load(open(GOLDEN / "replicate_2rank.json"))
    for r in range(2):
        assert json.load(open(tmp_path / f"r{r}.json")) == gold[str(r)]["tensors"], f"rank {r}"
