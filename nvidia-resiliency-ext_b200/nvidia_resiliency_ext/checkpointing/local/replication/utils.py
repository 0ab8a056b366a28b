"""Small helpers for the replication package (mirror of reference ``local/replication/utils.py``)."""


def zip_strict(*iterables):
    """``zip`` that insists on equal lengths and reports the lengths when they differ."""
    columns = [list(it) for it in iterables]
    sizes = [len(c) for c in columns]
    assert len(set(sizes)) <= 1, f"Tried to zip iterables of unequal lengths: {sizes}!"
    return zip(*columns)
