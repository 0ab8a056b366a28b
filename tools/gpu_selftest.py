"""Quick GPU bring-up check + micro-benchmark of the C-ABI hot path (run under gpurun).

Not a test-suite replacement: tests/ holds the parity tests.  This prints one table so a single GPU call
answers "is it correct" and "how fast is each walker" for several tile sizes.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nvidia-resiliency-ext_b200"))

import torch  # noqa: E402

from nvidia_resiliency_ext.checkpointing.b200 import _cabi  # noqa: E402
from nvidia_resiliency_ext.checkpointing.b200.engine import DeviceBuffer, Event, HostBuffer, Plan, Stream  # noqa: E402


def probe():
    for cmd in (
        "nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv",
        "nproc", "free -g | head -2", "df -h /dev/shm /tmp | cat", "ulimit -l", "cat /proc/sys/kernel/yama/ptrace_scope",
        "lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)|Thread'",
    ):
        try:
            print(f"$ {cmd}\n" + subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip())
        except Exception as e:  # noqa: BLE001
            print(f"$ {cmd} -> {e}")


def expected_pack(tensors, offsets, packed, total, narrow_mask):
    buf = torch.zeros(total, dtype=torch.uint8, device="cuda")
    for t, off, nb, nr in zip(tensors, offsets, packed, narrow_mask):
        if nb == 0:
            continue
        src = t.to(torch.bfloat16) if nr else t
        buf[off : off + nb] = src.contiguous().view(-1).view(torch.uint8)
    return buf


def ragged_set(gen):
    """Tensors with every alignment class: carved out of a flat byte buffer at odd offsets."""
    flat = torch.randint(0, 256, (64 << 20,), dtype=torch.uint8, device="cuda", generator=gen)
    out, cur = [], 0
    sizes = [0, 1, 3, 4, 15, 16, 17, 31, 32, 33, 63, 100, 255, 1000, 4096, 4097, 32768, 32769, 65536 + 5, 1 << 20, (1 << 20) + 13, 5 << 20]
    for mis in range(0, 16):
        for s in sizes:
            cur = (cur + 15) // 16 * 16 + mis
            out.append(flat[cur : cur + s])
            cur += s
    return flat, out


def check_correctness():
    gen = torch.Generator(device="cuda").manual_seed(1)
    flat, tensors = ragged_set(gen)
    f32 = []
    base = torch.randn(3_000_000, device="cuda", generator=gen)
    base[:8] = torch.tensor([float("inf"), float("-inf"), float("nan"), -0.0, 1e-40, -1e-45, 3.3895314e38, 1.0], device="cuda")
    cur = 0
    for mis in (0, 1, 2, 3):
        for n in (0, 1, 7, 8, 9, 1000, 8192, 8193, 300_001):
            cur = (cur + 3) // 4 * 4 + mis
            f32.append(base[cur : cur + n])
            cur += n
    ok = True
    for variant, vname in ((1, "ldg"), (2, "tma")):
        for tile in (4096, 32768, 65536):
            # bit-copy plan over byte tensors + fp32 tensors
            ts = tensors + f32
            ptrs = [t.data_ptr() if t.numel() else 0 for t in ts]
            nbytes = [t.numel() * t.element_size() for t in ts]
            plan = Plan(ptrs, nbytes, None, device=0, tile_bytes=tile, variant=variant)
            got = torch.zeros(max(plan.staging_bytes, 512), dtype=torch.uint8, device="cuda")[: plan.staging_bytes]
            stg_ptr = got.data_ptr()
            st = torch.cuda.current_stream().cuda_stream
            plan.pack(stg_ptr, st)
            torch.cuda.synchronize()
            exp = expected_pack(ts, plan.offsets, plan.packed_nbytes, plan.staging_bytes, [False] * len(ts))
            same = torch.equal(got, exp)
            # scatter back into fresh, differently aligned destinations
            dst_flat = torch.zeros(sum(nbytes) + 32 * len(ts) + 64, dtype=torch.uint8, device="cuda")
            dsts, c = [], 0
            for i, nb in enumerate(nbytes):
                c = (c + 15) // 16 * 16 + (i * 7) % 16
                dsts.append(dst_flat[c : c + nb])
                c += nb
            plan.update_ptrs([d.data_ptr() if d.numel() else 0 for d in dsts])
            plan.scatter(stg_ptr, st)
            torch.cuda.synchronize()
            rt = all(torch.equal(d, t.contiguous().view(-1).view(torch.uint8)) for d, t in zip(dsts, ts))
            print(f"copy   variant={vname} tile={tile:6d} tiles={plan.n_tiles:6d} pack_ok={same} scatter_ok={rt}")
            ok &= same and rt
            plan.close()
    # narrow: only fp32 tensors (4-byte aligned ones)
    al = [t for t in f32 if t.data_ptr() % 4 == 0]
    for tile in (4096, 65536):
        ptrs = [t.data_ptr() if t.numel() else 0 for t in al]
        nbytes = [t.numel() * 4 for t in al]
        plan = Plan(ptrs, nbytes, [1] * len(al), device=0, tile_bytes=tile)
        got = torch.zeros(max(plan.staging_bytes, 512), dtype=torch.uint8, device="cuda")[: plan.staging_bytes]
        stg_ptr = got.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        plan.pack(stg_ptr, st)
        torch.cuda.synchronize()
        exp = expected_pack(al, plan.offsets, plan.packed_nbytes, plan.staging_bytes, [True] * len(al))
        same = torch.equal(got, exp)
        outs = [torch.zeros_like(t) for t in al]
        plan.update_ptrs([o.data_ptr() if o.numel() else 0 for o in outs])
        plan.scatter(stg_ptr, st)
        torch.cuda.synchronize()
        wide = all(torch.equal(o.view(torch.int32), t.to(torch.bfloat16).to(torch.float32).view(torch.int32)) for o, t in zip(outs, al))
        print(f"narrow tile={tile:6d} pack_ok={same} widen_ok={wide}")
        ok &= same and wide
        plan.close()
    return ok


def llama_shard_tensors(gb, gen):
    """Llama-3-8B row-shard shapes x4 copies, truncated/scaled to ~gb GB (bench.py holds the exact C2 config)."""
    layer = [(512, 4096), (128, 4096), (128, 4096), (512, 4096), (1792, 4096), (1792, 4096), (512, 14336), (4096,), (4096,)]
    shapes = [(16032, 4096)] + layer * 32 + [(4096,), (16032, 4096)]
    tensors, total = [], 0
    for copy in range(4):
        for s in shapes:
            n = 1
            for d in s:
                n *= d
            if total + n * 4 > gb * (1 << 30):
                continue
            tensors.append(torch.empty(s, dtype=torch.float32, device="cuda").normal_(generator=gen))
            total += n * 4
    steps = [torch.full((), float(i), device="cuda") for i in range(291)]
    return tensors + steps, total


def bench(gb, iters):
    gen = torch.Generator(device="cuda").manual_seed(2)
    tensors, total = llama_shard_tensors(gb, gen)
    ptrs = [t.data_ptr() for t in tensors]
    nbytes = [t.numel() * 4 for t in tensors]
    print(f"bench state: {len(tensors)} tensors, {total/1e9:.3f} GB")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    stg = None

    def run(narrow, variant, vname, tile, env):
        nonlocal stg
        for k in ("NVRX_B200_TMA_STAGES", "NVRX_B200_TMA_CTAS_PER_SM", "NVRX_B200_LDG_CTAS_PER_SM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        flags = [1 if (narrow and t.numel() > 1) else 0 for t in tensors]
        try:
            plan = Plan(ptrs, nbytes, flags, device=0, tile_bytes=tile, variant=variant)
            if stg is None or stg.nbytes < plan.staging_bytes:
                stg = DeviceBuffer(plan.staging_bytes, 0)
            for direction in ("pack", "scatter"):
                fn = plan.pack if direction == "pack" else plan.scatter
                for _ in range(3):
                    fn(stg.ptr, st)
                e0, e1 = Event(0, True), Event(0, True)
                times = []
                for _ in range(iters):
                    e0.record(st)
                    fn(stg.ptr, st)
                    e1.record(st)
                    e1.synchronize()
                    times.append(e0.elapsed_ms(e1))
                times.sort()
                med = times[len(times) // 2]
                gbs = plan.algorithmic_bytes / med / 1e6
                rows.append((direction, vname, narrow, tile, env, med, gbs))
                print(f"{direction:7s} {vname} narrow={int(narrow)} tile={tile:6d} {env} med={med:8.3f} ms min={times[0]:8.3f} ms algo={gbs:8.1f} GB/s", flush=True)
            plan.close()
        except Exception as e:  # noqa: BLE001
            print(f"FAILED {vname} narrow={narrow} tile={tile} {env}: {e}", flush=True)

    for tile in (16384, 32768, 65536):
        for c in ("2", "4", "8"):
            run(False, 1, "ldg", tile, {"NVRX_B200_LDG_CTAS_PER_SM": c})
    for tile, stages, ctas in ((8192, "12", "1"), (8192, "12", "2"), (16384, "12", "1"), (16384, "6", "2"), (16384, "4", "3"),
                               (32768, "6", "1"), (32768, "3", "2"), (32768, "4", "1"), (65536, "3", "1")):
        run(False, 2, "tma", tile, {"NVRX_B200_TMA_STAGES": stages, "NVRX_B200_TMA_CTAS_PER_SM": ctas})
    for tile in (16384, 32768, 65536):
        for c in ("4", "8"):
            run(True, 1, "ldg", tile, {"NVRX_B200_LDG_CTAS_PER_SM": c})
    for k in ("NVRX_B200_TMA_STAGES", "NVRX_B200_TMA_CTAS_PER_SM", "NVRX_B200_LDG_CTAS_PER_SM"):
        os.environ.pop(k, None)
    # reference point: torch copy_ of the same bytes (read+write)
    a = torch.empty(total // 4, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record(); b.copy_(a); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"torch copy_ same bytes: med={ts[len(ts)//2]:.3f} ms -> {2*total/ts[len(ts)//2]/1e6:.1f} GB/s")
    del a, b
    # drain speed: staging -> pinned shm host buffer
    t0 = time.time()
    hb = HostBuffer.create(total, name=f"/nvrx_selftest_{os.getpid()}", pin=True, device=0, prefault_threads=16)
    print(f"hostbuf create+prefault+pin {total/1e9:.2f} GB: {time.time()-t0:.2f} s")
    side = Stream(0)
    lib = _cabi.lib()
    ev0, ev1 = Event(0, True), Event(0, True)
    for chunk in (0, 64 << 20, 256 << 20):
        for _ in range(2):
            ev0.record(side.handle)
            _cabi.check(lib.nvrx_drain(hb.data_ptr, stg.ptr, total, chunk, hb.progress_ptr, 0, side.handle, ev1.handle), "drain")
            ev1.synchronize()
            ms = ev0.elapsed_ms(ev1)
        print(f"drain chunk={chunk>>20:4d} MiB: {ms:.1f} ms -> {total/ms/1e6:.1f} GB/s (progress={hb.progress})")
    t0 = time.time()
    crc = hb.crc32(0, total, 16)
    print(f"crc32 16 threads: {time.time()-t0:.2f} s ({total/(time.time()-t0)/1e9:.1f} GB/s) crc={crc:#x}")
    path = "/dev/shm/nvrx_selftest.bin"
    fd = os.open(path, os.O_CREAT | os.O_WRONLY | os.O_TRUNC, 0o600)
    t0 = time.time()
    hb.write_fd(0, total, fd, 0, 16)
    os.close(fd)
    print(f"write_fd /dev/shm 16 threads: {time.time()-t0:.2f} s ({total/(time.time()-t0)/1e9:.1f} GB/s)")
    os.unlink(path)
    hb.close()
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-probe", action="store_true")
    args = ap.parse_args()
    if not args.no_probe:
        probe()
    ok = check_correctness()
    print("CORRECTNESS", "PASS" if ok else "FAIL")
    rows = bench(args.gb, args.iters)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "selftest.json"), "w") as f:
        json.dump({"ok": ok, "rows": rows}, f)
    sys.exit(0 if ok else 1)
