"""Golden outputs of the reference's ``checkpointing.utils.diff`` (utils.py:124-182) on the cases of tests/_diff_cases.py.

    python tests/golden/make_utils_golden.py        (build container only: imports /root/reference/src)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
assert os.path.isdir(REF_SRC), "the reference tree is needed to (re)generate golden vectors"
sys.path.insert(0, REF_SRC)  # the REFERENCE package, not the mirror
sys.path.insert(0, os.path.dirname(HERE))

from _diff_cases import cases, normal  # noqa: E402
from nvidia_resiliency_ext.checkpointing.utils import diff  # noqa: E402

import nvidia_resiliency_ext  # noqa: E402

assert nvidia_resiliency_ext.__file__.startswith(REF_SRC), nvidia_resiliency_ext.__file__
out = {}
for name, (left, right) in cases().items():
    try:
        out[name] = normal(diff(left, right))
    except Exception as exc:  # noqa: BLE001 - what the reference raises is part of its behaviour
        out[name] = {"raises": type(exc).__name__}
with open(os.path.join(HERE, "utils_diff.json"), "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
print(json.dumps(out, indent=1)[:3000])
