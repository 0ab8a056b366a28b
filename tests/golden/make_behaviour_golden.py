"""Outcomes (return values, exception types and messages) of the scenarios in tests/_error_cases.py when they run against the
REFERENCE package, one-rank gloo world, CPU only.

    python tests/golden/make_behaviour_golden.py        (build container only: imports /root/reference/src)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
sys.path.insert(0, REF_SRC)  # (also in the spawned checkpoint worker, which imports this file as its main module)
sys.path.insert(0, os.path.dirname(HERE))

if __name__ == "__main__":
    assert os.path.isdir(REF_SRC), "the reference tree is needed to (re)generate golden vectors"
    import torch.distributed as dist

    import nvidia_resiliency_ext

    assert nvidia_resiliency_ext.__file__.startswith(REF_SRC), nvidia_resiliency_ext.__file__
    from _error_cases import scenarios

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29671")
    dist.init_process_group("gloo", rank=0, world_size=1)
    with tempfile.TemporaryDirectory() as tmp:
        out = scenarios(tmp)
    dist.destroy_process_group()
    with open(os.path.join(HERE, "behaviour.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
