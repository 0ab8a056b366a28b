"""Inputs of the ``checkpointing.utils.diff`` cases (shared by tests/golden/make_utils_golden.py, which runs the REFERENCE on
them, and tests/test_oracle_golden.py, which runs the mirror)."""
import numpy as np
import torch


class Sharded:
    """Anything with ``replica_id`` and ``data`` (the reference's duck-typed branch, utils.py:167-172)."""

    def __init__(self, data, replica_id=0):
        self.data, self.replica_id = data, replica_id


class Other(Sharded):
    pass


def cases():
    t = torch.arange(6.0)
    return {
        "equal_nested": ({"a": {"w": t.clone(), "n": 3}, "l": [1, "x", t.clone()]}, {"a": {"w": t.clone(), "n": 3}, "l": [1, "x", t.clone()]}),
        "missing_keys_both_sides": ({"a": 1, "b": {"c": 2, "d": 3}}, {"b": {"c": 2, "e": 4}, "z": 0}),
        "tensor_value_mismatch": ({"w": t.clone(), "v": t.clone()}, {"w": t + 1, "v": t.clone()}),
        "tensor_vs_scalar": ({"w": t.clone()}, {"w": 5}),
        "dtype_differs_values_equal": ({"w": t.clone()}, {"w": t.to(torch.float64)}),
        "list_longer_left": ([1, 2, 3, 4], [1, 2]),
        "list_longer_right": ([1], [1, 5, 6]),
        "list_element_mismatch": ([1, {"k": t.clone()}, 3], [1, {"k": t * 2}, 4]),
        "tuple_nested": ((1, (2, 3)), (1, (2, 4))),
        "ndarray_elements": (np.array([1, 2, 3]), np.array([1, 0, 3])),
        "replica_id_objects": ({"s": Sharded({"x": t.clone(), "y": 1})}, {"s": Sharded({"x": t + 1, "z": 1}, replica_id=1)}),
        "incomparable_shapes": ({"w": torch.zeros(3)}, {"w": torch.zeros(4)}),
        "strings_and_none": ({"a": "x", "b": None, "c": 1.5}, {"a": "y", "b": None, "c": 1.5}),
        "prefix_types": ({1: {"a": 0}, (2, 3): {"a": 0}}, {1: {"a": 1}, (2, 3): {"b": 0}}),
    }


def normal(result):
    """JSON-able, order-insensitive form of ``(only_left, only_right, mismatch)``: keys of a dict come out of set operations
    in an unspecified order in the reference (``x1.keys() - x2.keys()``), so each list is sorted by its repr."""
    def pre(p):
        if isinstance(p, tuple):
            return [pre(x) for x in p]
        if isinstance(p, type):
            return f"<{p.__name__}>"
        return p

    only_left, only_right, mismatch = result
    return {
        "only_left": sorted((pre(p) for p in only_left), key=repr),
        "only_right": sorted((pre(p) for p in only_right), key=repr),
        "mismatch": sorted(([pre(p), a.__name__, b.__name__] for p, a, b in mismatch), key=repr),
    }
