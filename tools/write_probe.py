"""Host-only probe (no GPU work): how fast can a 16 GB slot be persisted into a tmpfs file on this box?
pwrite pool vs mmap+memcpy pool, thread counts, fresh vs pre-existing file pages."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nvidia-resiliency-ext_b200"))
from nvidia_resiliency_ext.checkpointing.b200.engine import HostBuffer  # noqa: E402

n = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 8 << 30
hb = HostBuffer.create(n, pin=False, prefault_threads=16)
path = "/dev/shm/nvrx_write_probe.bin"
for mode in ("pwrite", "mmap"):
    for threads in (4, 16, 32):
        os.environ.pop("NVRX_B200_WRITE_PWRITE", None)
        if mode == "pwrite":
            os.environ["NVRX_B200_WRITE_PWRITE"] = "1"
        fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
        os.ftruncate(fd, n)
        t = time.time()
        hb.write_fd(0, n, fd, 0, threads)
        fresh = time.time() - t
        t = time.time()
        hb.write_fd(0, n, fd, 0, threads)  # pages exist now
        again = time.time() - t
        os.close(fd)
        os.unlink(path)
        print(f"{mode:6s} threads={threads:2d}: fresh file {n/fresh/1e9:6.2f} GB/s, existing pages {n/again/1e9:6.2f} GB/s", flush=True)
hb.close()
