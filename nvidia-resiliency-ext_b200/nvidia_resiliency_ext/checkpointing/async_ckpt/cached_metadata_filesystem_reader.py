"""``FileSystemReader`` that reads a checkpoint's ``.metadata`` once per process.

Same role as reference ``checkpointing/async_ckpt/cached_metadata_filesystem_reader.py`` (``CachedMetadataFileSystemReader``
``:24``): several loads of one checkpoint directory (model, optimizer, ... loaded by separate ``dcp.load`` calls) unpickle the
global metadata only the first time.  The cache is per class, keyed by the absolute checkpoint path."""

import os
from typing import Dict, Optional, Union

from torch.distributed.checkpoint import FileSystemReader, Metadata


class CachedMetadataFileSystemReader(FileSystemReader):
    _metadata_cache: Dict[str, Metadata] = {}

    def __init__(self, path: Union[str, os.PathLike], cache_metadata: bool = True) -> None:
        """``cache_metadata=False`` makes this a plain ``FileSystemReader`` (every ``read_metadata`` hits the disk)."""
        super().__init__(path=path)
        self._cache_key: Optional[str] = os.path.abspath(os.fspath(path)) if cache_metadata else None

    def read_metadata(self) -> Metadata:
        key = self._cache_key
        if key is None:
            return super().read_metadata()
        cache = type(self)._metadata_cache
        if key not in cache:
            cache[key] = super().read_metadata()
        return cache[key]

    @classmethod
    def clear_metadata_cache(cls, path: Union[str, os.PathLike, None] = None) -> None:
        """Forget one checkpoint directory (e.g. after it was overwritten), or everything when ``path`` is None."""
        if path is None:
            cls._metadata_cache.clear()
        else:
            cls._metadata_cache.pop(os.path.abspath(os.fspath(path)), None)
