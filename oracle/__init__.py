"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithm for the checkpoint-snapshot path (``snapshot_oracle.py``: flattening, per-tensor
snapshot, packed layout, bf16 rounding, clique / coverage logic, the reference save path as a timed baseline) and of what the
GPU checksum path must produce (``crc_oracle.py``: zlib), plus one C++ helper that runs the checksum kernel's own device
functions on the host (``crc_lanes.cpp``, built by ``oracle/Makefile`` into ``oracle/_build``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import or
execute anything in here -- as the checker or the baseline, never as the thing measured or shipped.  The product
(``nvidia-resiliency-ext_b200/``) has no CPU fallback and does not import this package."""
