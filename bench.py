#!/usr/bin/env python
"""Benchmark of the checkpoint-snapshot hot path (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own path (oracle port), host cores

Workload (N=1 and every N: weak scaling, no data-path collective): BASELINE config C2 -- per rank a 1/8 row
shard of Llama-3-8B in 4 fp32 copies (param, main_param, exp_avg, exp_avg_sq) = 1164 tensors, 16,060,522,496 B,
plus 291 scalar fp32 `step` tensors, synthetic values.  ``--narrow`` switches to C3.

One "step" = one snapshot of the whole state dict.
  value      device snapshot rate: S bytes / pack-kernel time (state resident in HBM; inputs >> L2 so no flush),
             CUDA events on the launching stream, K back-to-back launches, max over ranks, summed over ranks.
  roofline   algorithmic HBM bytes of one pack launch (2*S, or 1.5*S_in when narrowing) / that time, against the
             measured copy peak in MEASURED_PEAKS.json.
  e2e        same metric through the public API (TorchAsyncCheckpoint.async_save on the GPU state dict) until the
             snapshot bytes are in host memory: includes the D2H of the whole snapshot every step.
  stall_ms   how long the training stream is held up by async_save (call + pack kernel); persist_s: until
             finalize_async_save(blocking=True) returned (file on /dev/shm loadable by torch.load).
  cpu_baseline / --impl reference: the reference's save path (per-tensor pinned D2H + cuda sync + fork + torch.save),
             oracle port, timed on this box's host cores.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nvidia-resiliency-ext_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "checkpoint_snapshot_GBps"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------
def llama3_8b_shard_shapes():
    """1/8 row shard (dim0 / 8) of every Llama-3-8B tensor (vocab 128256, hidden 4096, 32 layers, ffn 14336, 32/8
    heads): 291 tensors, 1,003,782,656 parameters per rank."""
    shapes = [("embed", (16032, 4096))]
    for layer in range(32):
        for name, shp in (("q", (512, 4096)), ("k", (128, 4096)), ("v", (128, 4096)), ("o", (512, 4096)),
                          ("gate", (1792, 4096)), ("up", (1792, 4096)), ("down", (512, 14336)), ("ln1", (512,)), ("ln2", (512,))):
            shapes.append((f"layers.{layer}.{name}", shp))
    shapes += [("final_ln", (512,)), ("lm_head", (16032, 4096))]
    return shapes


def llama3_8b_shard_state(device, seed=1234, scale=1.0):
    """{"model": {name: p}, "optimizer": {"state": {i: {main_param, exp_avg, exp_avg_sq, step}}}}."""
    g = torch.Generator(device=device).manual_seed(seed)
    model, opt = {}, {}
    total = 0
    for i, (name, shp) in enumerate(llama3_8b_shard_shapes()):
        if scale != 1.0 and len(shp) == 2:
            shp = (max(1, int(shp[0] * scale)), shp[1])
        p = torch.empty(shp, dtype=torch.float32, device=device).normal_(0, 0.02, generator=g)
        m = p.clone()
        ea = torch.empty(shp, dtype=torch.float32, device=device).normal_(0, 1e-3, generator=g)
        es = torch.empty(shp, dtype=torch.float32, device=device).uniform_(0, 1e-6, generator=g)
        if i == 1:  # edge values for the narrowing path
            flat = ea.view(-1)
            flat[:6] = torch.tensor([float("inf"), float("-inf"), float("nan"), -0.0, 1e-40, -1e-45], device=device)
        model[name] = p
        opt[i] = {"main_param": m, "exp_avg": ea, "exp_avg_sq": es, "step": torch.tensor(float(1000 + i), device=device)}
        total += 4 * p.numel() * 4 + 4
    return {"model": model, "optimizer": {"state": opt}}, total


def flatten(sd):
    out = []

    def walk(x):
        for v in (x.values() if isinstance(x, dict) else x):
            if isinstance(v, (dict, list)):
                walk(v)
            elif isinstance(v, torch.Tensor):
                out.append(v)

    walk(sd)
    return out


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        if shutil.which("nvidia-smi"):
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append(ln)

            threading.Thread(target=pump, daemon=True).start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:  # nvidia-smi needs a moment (longer on 8-GPU boxes)
                time.sleep(0.05)
        return self

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(label_prefix):
    """dram read+write bytes per launch of the dominant kernel from the committed `ncu --set full` capture."""
    p = ROOT / "profiles" / "r01_ncu_walk_summary.json"
    try:
        for k in json.load(open(p))["kernels"]:
            if k["label"].startswith(label_prefix):
                return float(k["traffic_bytes"])
    except Exception:  # noqa: BLE001
        pass
    return None


def init_dist(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return rank, world, local


def max_over_ranks(x: float) -> float:
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2]


def shm_dir(rank):
    base = Path("/dev/shm") if os.access("/dev/shm", os.W_OK) else Path("/tmp")
    d = base / f"nvrx_b200_bench_{os.getpid()}_{rank}"
    d.mkdir(parents=True, exist_ok=True)
    return d


# ----------------------------------------------------------------------------------------------------
# arms
# ----------------------------------------------------------------------------------------------------
def run_engine(args, rank, world, local):
    from nvidia_resiliency_ext.checkpointing.async_ckpt.torch_ckpt import TorchAsyncCheckpoint
    from nvidia_resiliency_ext.checkpointing.b200.engine import Event, SnapshotEngine

    dev = torch.device("cuda", local)
    sd, total = llama3_8b_shard_state(dev, seed=1234 + rank, scale=args.scale)
    tensors = flatten(sd)
    engine = SnapshotEngine.get(local, host_slots=2)  # slots are created lazily; sequential steps use one
    stream = torch.cuda.current_stream().cuda_stream

    # ---- kernel-only: K back-to-back pack launches (value, roofline) -------------------------------
    mask = SnapshotEngine._narrow_mask(tensors, args.narrow)
    plan = engine._plan_for(tensors, mask)
    staging = engine._ensure_staging(plan.staging_bytes)
    clocks = ClockSampler(local)
    clocks.__enter__()  # sampled from here until the end of the end-to-end loop (both timed regions)
    for _ in range(args.warmup):
        plan.pack(staging.ptr, stream)
    e0, e1 = Event(local, True), Event(local, True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record(stream)
    for _ in range(args.steps):
        plan.pack(staging.ptr, stream)
    e1.record(stream)
    e1.synchronize()
    kernel_ms = max_over_ranks(e0.elapsed_ms(e1) / args.steps)
    torch.cuda.synchronize()
    dist.barrier()
    launches = args.steps

    # ---- end to end through the public API -----------------------------------------------------------
    # Host-memory budget: the box's cgroup (200 GiB on the 1-GPU boxes of this pool) also counts tmpfs / POSIX-shm
    # pages.  One rank needs a 16 GB pinned slot; persisting adds a 16 GB file per rank.  With several ranks on a box
    # the files would not fit (8 x 32 GB), so N>1 measures the snapshot through utils.preload_tensors (state -> pinned
    # host, no file) and only N=1 runs the full TorchAsyncCheckpoint.async_save -> file -> restore cycle.
    persist_files = world == 1 or args.persist
    out_dir = shm_dir(rank)
    engine_launches_before = None
    ckpt = TorchAsyncCheckpoint(persistent_queue=True, narrow_fp32_to_bf16=args.narrow) if persist_files else None
    stall, host_safe, persist = [], [], []
    done_ev = torch.cuda.Event()
    path = out_dir / "ckpt.pt"  # one file per rank, overwritten every step
    from nvidia_resiliency_ext.checkpointing.utils import preload_tensors

    for it in range(args.e2e_warmup + args.e2e_steps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if persist_files:
            ckpt.async_save(sd, path)
            done_ev.record()
            done_ev.synchronize()  # the training stream is free again here
            t1 = time.perf_counter()
            snap = next(iter(ckpt._pending.values()))
            snap.wait()  # bytes are in host memory
            t2 = time.perf_counter()
            ckpt.finalize_async_save(blocking=True, no_dist=True)
            t3 = time.perf_counter()
        else:
            _, snap = preload_tensors(sd, non_blocking=True, narrow=args.narrow, return_snapshot=True)
            done_ev.record()
            done_ev.synchronize()
            t1 = time.perf_counter()
            snap.wait()
            t2 = t3 = time.perf_counter()
            snap.release()
        if it == args.e2e_warmup - 1 or (args.e2e_warmup == 0 and it == 0 and engine_launches_before is None):
            engine_launches_before = engine.launches if args.e2e_warmup else 0
        if it >= args.e2e_warmup:
            stall.append(t1 - t0)
            host_safe.append(t2 - t0)
            persist.append(t3 - t0)
    clocks.__exit__(None, None, None)
    launches += engine.launches - (engine_launches_before or 0)  # pack sub-launches of the pipelined snapshots
    e2e_s = max_over_ranks(median(host_safe))
    stall_ms = max_over_ranks(median(stall)) * 1e3
    persist_s = max_over_ranks(median(persist))
    linked = bool(persist_files and path.exists() and os.stat(path).st_nlink > 1)  # the file IS the pinned slot (opt-in mode)
    # spot-check the last file against the live state (bit-exact unless narrowed)
    check = "skipped"
    if rank == 0 and not args.no_verify and persist_files:
        loaded = torch.load(path, weights_only=False)
        lt = flatten(loaded)
        ok = len(lt) == len(tensors)
        for a, b in list(zip(lt, tensors))[:: max(1, len(tensors) // 64)]:
            want = b.to(torch.bfloat16).cpu() if (args.narrow and b.dtype == torch.float32 and b.numel() > 0) else b.cpu()
            ok &= a.dtype == want.dtype and a.shape == want.shape and (
                a.numel() == 0 or torch.equal(a.contiguous().view(-1).view(torch.uint8), want.contiguous().view(-1).view(torch.uint8)))
        check = "bit-exact" if ok else "MISMATCH"
        del loaded, lt
    # restore (C5 on one GPU): file -> mmap -> parallel gather into the pinned slot -> ONE H2D -> ONE scatter kernel
    restore_s = None
    if not args.no_restore and persist_files:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        loaded = torch.load(path, weights_only=False, mmap=True)
        lt = flatten(loaded)
        widen = [torch.float32 if (args.narrow and t.dtype == torch.bfloat16) else None for t in lt]
        resident = None
        if os.environ.get("NVRX_B200_ZERO_COPY") == "1":  # opt-in: the file is a hard link to a slot that is still pinned
            resident = engine.resident_source(path, lt)
        file_source = None
        if resident is None and os.environ.get("NVRX_B200_RESTORE_PREAD") == "1":  # opt-in A/B: pread instead of mmap gather
            from nvidia_resiliency_ext.checkpointing.b200.ptzip import tensor_offsets_in_file

            offs = tensor_offsets_in_file(path, lt)
            file_source = (str(path), offs) if offs is not None else None
        back = engine.restore(lt, widen_to=widen if args.narrow else None, resident=resident, file_source=file_source)
        torch.cuda.synchronize()
        restore_s = max_over_ranks(time.perf_counter() - t0)
        launches += 1
        if not args.narrow:
            for a, b in list(zip(back, tensors))[:: max(1, len(tensors) // 32)]:
                if not torch.equal(a.view(-1).view(torch.uint8), b.view(-1).view(torch.uint8)):
                    check = "MISMATCH(restore)"
        del loaded, lt, back
    if ckpt is not None:
        ckpt.close()
    shutil.rmtree(out_dir, ignore_errors=True)

    packed_bytes = plan.staging_bytes
    algo = plan.algorithmic_bytes
    peak, peak_src = hbm_peak()
    achieved = algo / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC,
        "value": round(world * total / (kernel_ms * 1e-3) / 1e9, 1),
        "unit": "GB/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": round(kernel_ms, 4),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16-narrowed bytes" if args.narrow else "u8",
        "data": "synthetic",
        "config": {
            "workload": ("C3" if args.narrow else "C2") + ": 16.06 GB Llama-3-8B-shaped param+Adam state_dict per rank "
                        "(1164 fp32 tensors + 291 scalar steps), pack + D2H drain",
            "state_bytes_per_rank": total,
            "packed_bytes_per_rank": packed_bytes,
            "tensors": len(tensors),
            "l2": "inputs (16 GB) >> L2 (126 MB); no flush needed",
            "walker": os.environ.get("NVRX_B200_VARIANT", "auto"),
            "scale": args.scale,
        },
        "e2e": {
            "value": round(world * total / e2e_s / 1e9, 2),
            "unit": "GB/s",
            "h2d_bytes_per_step": int(len(tensors) * 32),
            "d2h_bytes_per_step": int(packed_bytes),
            "definition": ("state bytes / wall time from TorchAsyncCheckpoint.async_save() until the snapshot is in pinned host memory"
                           if persist_files else
                           "state bytes / wall time from checkpointing.utils.preload_tensors() until the snapshot is in pinned host memory "
                           "(no file at N>1: host-memory budget of the box)"),
            "steps": args.e2e_steps,
        },
        "stall_ms": round(stall_ms, 3),
        "persist_s": round(persist_s, 3) if persist_files else None,
        "persist_GBps": round(world * total / persist_s / 1e9, 2) if persist_files else None,
        "persist_mode": (("zero-copy link" if linked else "parallel copy") if persist_files else None),
        "restore_s": None if restore_s is None else round(restore_s, 3),
        "restore_GBps": None if restore_s is None else round(world * total / restore_s / 1e9, 2),
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": args.traffic_bytes or ncu_traffic("pack LDG narrow" if args.narrow else "pack TMA"),
            "traffic_source": "profiles/r01_ncu_walk_summary.json (ncu --set full, same state, one launch)", "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
            "kernel": "nvrx::walk_tma<pack>" if not args.narrow else "nvrx::walk_ldg<pack> (narrow)",
        },
        "clocks": clocks.summary(),
        "verify": check,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(sd, total, out_dir.parent, sample_gb=args.baseline_sample_gb)
    return line


def cpu_baseline(sd, total, base_dir, sample_gb):
    """The reference save path (oracle port) on this box: D2H of the FULL state (stall), persistence on a bounded
    sample of the tensors."""
    from oracle import snapshot_oracle as orc

    tensors = flatten(sd)
    # warm the pinned-memory cache like a second checkpoint would see it
    orc.reference_preload(sd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pre = orc.reference_preload(sd)
    torch.cuda.synchronize()
    d2h = time.perf_counter() - t0
    del pre
    sample, acc = {}, 0
    for i, t in enumerate(tensors):
        if acc + t.numel() * 4 > sample_gb * 1e9:
            continue
        sample[f"t{i}"] = t
        acc += t.numel() * 4
    path = Path(base_dir) / f"nvrx_ref_sample_{os.getpid()}.pt"
    res = orc.reference_fork_save(sample, path)
    path.unlink(missing_ok=True)
    persist_rate = acc / (res["total"] - res["d2h"]) / 1e9
    return {
        "value": round(total / d2h / 1e9, 2), "unit": "GB/s", "cores": 1, "kind": "port",
        "sample": f"full-state per-tensor pinned D2H + cuda sync (stall {d2h*1e3:.0f} ms); fork + torch.save timed on "
                  f"{acc/1e9:.2f} GB of the tensors ({persist_rate:.2f} GB/s, 1 core)",
        "stall_ms": round(d2h * 1e3, 1), "persist_GBps_1core": round(persist_rate, 2), "host_cores_on_box": os.cpu_count(),
    }


def run_reference(args, rank, world, local):
    """The reference's own path (oracle port of utils.preload_tensors + torch_ckpt.async_save + fork/torch.save)."""
    from oracle import snapshot_oracle as orc

    if rank != 0:
        return None
    dev = torch.device("cuda", local)
    sd, total = llama3_8b_shard_state(dev, seed=1234, scale=args.scale)
    tensors = flatten(sd)
    out_dir = shm_dir(rank)
    times = []
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pre = orc.reference_preload(sd)  # per-tensor .to("cpu", non_blocking=True)  (utils.py:85-99)
        torch.cuda.synchronize()         # torch_ckpt.py:50
        dt = time.perf_counter() - t0
        del pre
        if it >= args.warmup:
            times.append(dt)
    d2h = median(times)
    sample, acc = {}, 0
    for i, t in enumerate(tensors):
        if acc + t.numel() * 4 > args.baseline_sample_gb * 1e9:
            continue
        sample[f"t{i}"] = t
        acc += t.numel() * 4
    res = orc.reference_fork_save(sample, out_dir / "ref.pt")
    shutil.rmtree(out_dir, ignore_errors=True)
    persist_rate = acc / (res["total"] - res["d2h"]) / 1e9
    value = round(total / d2h / 1e9, 2)
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(d2h * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: 16.06 GB Llama-3-8B-shaped param+Adam state_dict per rank (1164 fp32 tensors + 291 scalar steps), "
                               "reference path: per-tensor pinned D2H + torch.cuda.synchronize()", "state_bytes_per_rank": total,
                   "tensors": len(tensors), "scale": args.scale},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stall_ms": round(d2h * 1e3, 2),
        "persist_GBps": round(persist_rate, 2),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": f"every step: full-state preload_tensors + cuda sync; once: fork + torch.save of {acc/1e9:.2f} GB "
                                   f"({persist_rate:.2f} GB/s on 1 of {os.cpu_count()} host cores)"},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--narrow", action="store_true", help="C3: fp32->bf16 in the pack kernel")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the state (debug only; numbers at != 1.0 are not C2)")
    ap.add_argument("--baseline-sample-gb", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-restore", action="store_true")
    ap.add_argument("--persist", action="store_true", help="write checkpoint files at N>1 too (needs ~32 GB of host memory per rank)")
    ap.add_argument("--traffic-bytes", type=float, default=None, help="dram read+write bytes per launch from the ncu capture")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else max(args.warmup, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the snapshot path has no CPU fallback)")
    rank, world, local = init_dist(args.gpus)
    line = run_reference(args, rank, world, local) if args.impl == "reference" else run_engine(args, rank, world, local)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if args.impl == "engine":
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
