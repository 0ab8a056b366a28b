"""Public API surface of the reference's checkpointing package (names, parameters, defaults), dumped by IMPORTING THE REFERENCE:
tests/test_api_surface_cpu.py holds the mirror against it (same names, argument meaning, defaults).

    python tests/golden/make_api_golden.py        (build container only: imports /root/reference/src)
"""
import importlib
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
MODULES = [
    "nvidia_resiliency_ext.checkpointing.utils",
    "nvidia_resiliency_ext.checkpointing.async_ckpt.core",
    "nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt",
    "nvidia_resiliency_ext.checkpointing.async_ckpt.filesystem_async",
    "nvidia_resiliency_ext.checkpointing.async_ckpt.state_dict_saver",
    "nvidia_resiliency_ext.checkpointing.local.base_state_dict",
    "nvidia_resiliency_ext.checkpointing.local.basic_state_dict",
    "nvidia_resiliency_ext.checkpointing.local.ckpt_managers.base_manager",
    "nvidia_resiliency_ext.checkpointing.local.ckpt_managers.local_manager",
    "nvidia_resiliency_ext.checkpointing.local.replication.strategies",
    "nvidia_resiliency_ext.checkpointing.local.replication.group_utils",
    "nvidia_resiliency_ext.checkpointing.local.replication.torch_device_utils",
    "nvidia_resiliency_ext.checkpointing.local.replication.utils",
]


def params(fn):
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    out = []
    for p in sig.parameters.values():
        default = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append({"name": p.name, "kind": p.kind.name, "default": default})
    return out


def describe(module_name):
    mod = importlib.import_module(module_name)
    out = {}
    for name, obj in vars(mod).items():
        if name.startswith("_") or getattr(obj, "__module__", None) != module_name:
            continue
        if inspect.isclass(obj):
            members = {}
            for mname, raw in vars(obj).items():
                if mname.startswith("_") and mname != "__init__":
                    continue
                kind = "method"
                fn = raw
                if isinstance(raw, staticmethod):
                    kind, fn = "staticmethod", raw.__func__
                elif isinstance(raw, classmethod):
                    kind, fn = "classmethod", raw.__func__
                elif isinstance(raw, property):
                    members[mname] = {"kind": "property"}
                    continue
                if not callable(fn):
                    members[mname] = {"kind": "attribute"}
                    continue
                members[mname] = {"kind": kind, "params": params(fn), "abstract": bool(getattr(fn, "__isabstractmethod__", False))}
            out[name] = {"type": "class", "bases": [b.__name__ for b in obj.__mro__[1:] if b is not object], "members": members}
        elif inspect.isfunction(obj):
            out[name] = {"type": "function", "params": params(obj)}
    return out


def surface():
    return {m: describe(m) for m in MODULES}


if __name__ == "__main__":
    REF_SRC = "/root/reference/src"
    assert os.path.isdir(REF_SRC), "the reference tree is needed to (re)generate golden vectors"
    sys.path.insert(0, REF_SRC)
    import nvidia_resiliency_ext

    assert nvidia_resiliency_ext.__file__.startswith(REF_SRC), nvidia_resiliency_ext.__file__
    data = surface()
    with open(os.path.join(HERE, "api_surface.json"), "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    print({m: len(v) for m, v in data.items()})
