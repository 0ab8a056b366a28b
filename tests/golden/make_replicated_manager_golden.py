"""What the REFERENCE's LocalCheckpointManager + CliqueReplicationStrategy do in the flow of tests/_replicated_manager_flow.py
(4 gloo ranks, CPU tensors): files per node, find_latest, what load() returns after a node lost its directory, cleanup.

    python tests/golden/make_replicated_manager_golden.py        (build container only: imports /root/reference/src)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, world, port, base, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import nvidia_resiliency_ext

    assert nvidia_resiliency_ext.__file__.startswith(REF_SRC), nvidia_resiliency_ext.__file__
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the reference's isend / irecv_state_dict stage host tensors through the GPU (``ten.cuda()`` group_utils.py:394,
    # ``torch.empty_like(ten, device='cuda')`` :444); on this CPU-only box both are redirected to host memory -- shims on torch,
    # the reference's code is untouched
    import torch

    real_empty_like = torch.empty_like
    torch.empty_like = lambda t, *a, **k: real_empty_like(t, *a, **{**k, "device": "cpu"} if str(k.get("device", "")).startswith("cuda") else k)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from _replicated_manager_flow import run

    ret[rank] = run(rank, base)
    dist.destroy_process_group()


if __name__ == "__main__":
    assert os.path.isdir(REF_SRC), "the reference tree is needed to (re)generate golden vectors"
    import torch.multiprocessing as mp

    from _replicated_manager_flow import WORLD

    with mp.Manager() as m, tempfile.TemporaryDirectory() as base:
        ret = m.dict()
        mp.spawn(_worker, args=(WORLD, 29677, base, ret), nprocs=WORLD, join=True)
        out = {str(r): ret[r] for r in range(WORLD)}
    with open(os.path.join(HERE, "replicated_manager_4rank.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out["1"], indent=1)[:2500])
