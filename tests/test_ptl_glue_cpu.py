"""HierarchicalCheckpointIO / LocalCheckpointCallback routing (reference tests/ptl_resiliency/unit/
test_local_ckpt_callback.py uses an in-memory manager the same way).  lightning is not installed here: a minimal stub
of the three lightning symbols the module imports is injected."""
import sys
import types

import pytest


@pytest.fixture
def glue(monkeypatch):
    pl = types.ModuleType("lightning.pytorch")

    class ModelCheckpoint:
        def __init__(self, every_n_train_steps=None, train_time_interval=None):
            self.every_n_train_steps, self.train_time_interval = every_n_train_steps, train_time_interval

    class _WrappingCheckpointIO:
        def __init__(self, checkpoint_io=None):
            self.checkpoint_io = checkpoint_io

    pl.callbacks = types.SimpleNamespace(ModelCheckpoint=ModelCheckpoint)
    wrapper = types.ModuleType("lightning.pytorch.plugins.io.wrapper")
    wrapper._WrappingCheckpointIO = _WrappingCheckpointIO
    root = types.ModuleType("lightning")
    root.pytorch = pl
    root.__spec__ = types.SimpleNamespace(name="lightning")
    for name, mod in {"lightning": root, "lightning.pytorch": pl, "lightning.pytorch.plugins": types.ModuleType("p"),
                      "lightning.pytorch.plugins.io": types.ModuleType("io"), "lightning.pytorch.plugins.io.wrapper": wrapper}.items():
        monkeypatch.setitem(sys.modules, name, mod)
    import importlib.util

    real_find = importlib.util.find_spec
    monkeypatch.setattr(importlib.util, "find_spec", lambda n, *a: root.__spec__ if n == "lightning" else real_find(n, *a))
    sys.modules.pop("nvidia_resiliency_ext.ptl_resiliency.local_checkpoint_callback", None)
    import nvidia_resiliency_ext.ptl_resiliency.local_checkpoint_callback as mod

    yield mod
    sys.modules.pop("nvidia_resiliency_ext.ptl_resiliency.local_checkpoint_callback", None)


class MemoryManager:
    def __init__(self):
        self.saved, self.latest = [], -1

    def save(self, tasd, iteration, is_async=False):
        self.saved.append((tasd, iteration, is_async))
        self.latest = iteration
        return "request" if is_async else None

    def find_latest(self):
        return self.latest

    def load(self):
        return self.saved[-1][0], (self.latest, 0, "")


class GlobalIO:
    def __init__(self):
        self.calls = []

    def save_checkpoint(self, ckpt, path, storage_options=None):
        self.calls.append(("save", path))

    def load_checkpoint(self, path, map_location=None, **kw):
        self.calls.append(("load", path))
        return {"from": "global"}

    def remove_checkpoint(self, path):
        self.calls.append(("remove", path))


def test_routing(glue):
    class IO(glue.HierarchicalCheckpointIO):
        def to_tensor_aware_state_dict(self, checkpoint):
            return ("tasd", checkpoint)

        def from_tensor_aware_state_dict(self, tasd, **kw):
            return {"from": "local", "payload": tasd}

    mgr, gio = MemoryManager(), GlobalIO()
    make = IO.get_partial_wrapper_constructor(mgr, get_global_ckpt_iteration_fn=lambda p: int(str(p).split("=")[-1]))
    io = make(gio)
    io.async_save = True
    # global save: no local options
    io.save_checkpoint({"w": 1}, "/g/step=10")
    assert gio.calls == [("save", "/g/step=10")] and not mgr.saved
    # nothing local yet -> global load
    assert io.load_checkpoint("/g/step=10") == {"from": "global"}
    # local save through the callback's storage_options
    opts = {glue.LOCAL_CKPT_OPTS_KEY: dict(ckpt_type="local", iteration=12)}
    assert io.save_checkpoint({"w": 2}, None, storage_options=opts) == "request"
    assert mgr.saved == [(("tasd", {"w": 2}), 12, True)]
    with pytest.raises(ValueError):
        io.save_checkpoint({"w": 2}, "/some/path", storage_options=opts)
    opts[glue.LOCAL_CKPT_OPTS_KEY]["is_async"] = False
    opts[glue.LOCAL_CKPT_OPTS_KEY]["iteration"] = 14
    assert io.save_checkpoint({"w": 3}, None, storage_options=opts) is None
    # local (14) newer than global (10) -> resume locally; older than global (20) -> resume globally
    assert io.load_checkpoint("/g/step=10")["from"] == "local"
    assert io.load_checkpoint("/g/step=14")["from"] == "local"
    assert io.load_checkpoint("/g/step=20") == {"from": "global"}
    io.remove_checkpoint("/g/step=10")
    assert gio.calls[-1] == ("remove", "/g/step=10")


def test_callback_only_saves_local_last(glue):
    cb = glue.LocalCheckpointCallback(every_n_train_steps=20)
    assert cb.every_n_train_steps == 20
    saved = []
    trainer = type("T", (), {"global_step": 40, "save_checkpoint": lambda self, path, storage_options=None: saved.append((path, storage_options))})()
    cb.on_train_epoch_end(trainer, None)
    cb.on_validation_end(trainer, None)
    cb._save_topk_checkpoint(trainer, {})
    assert saved == []
    cb._save_last_checkpoint(trainer, {})
    assert saved == [(None, {glue.LOCAL_CKPT_OPTS_KEY: {"ckpt_type": "local", "iteration": 40}})]
