"""Caller-side demo of the asynchronous ``torch.distributed.checkpoint`` writer (the flow of reference
examples/checkpointing/async_writer.py): plan -> stage -> schedule -> train on -> finalize -> load back.

Works on CUDA (NCCL; the staging is one engine snapshot) and, for trying the API, on CPU (gloo; host tensors only):

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/async_writer_demo.py [--persistent] [--cache]
"""
import argparse
import os
import shutil
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "nvidia-resiliency-ext_b200"))

import torch
import torch.distributed as dist
import torch.distributed.checkpoint as dcp
import torch.nn as nn
from torch.distributed.checkpoint import DefaultSavePlanner, FileSystemReader

from nvidia_resiliency_ext.checkpointing.async_ckpt.core import AsyncCallsQueue, AsyncRequest
from nvidia_resiliency_ext.checkpointing.async_ckpt.filesystem_async import FileSystemWriterAsync
from nvidia_resiliency_ext.checkpointing.async_ckpt.state_dict_saver import (
    save_state_dict_async_finalize,
    save_state_dict_async_plan,
)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt_dir", default="/tmp/nvrx_b200_dcp_demo")
    ap.add_argument("--persistent", action="store_true", help="persistent (spawned) writer process instead of a fork per save")
    ap.add_argument("--cache", action="store_true", help="reuse the save plan while the state dict keeps its structure")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    dist.init_process_group("nccl" if cuda else "gloo")
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")

    model = nn.Sequential(nn.Linear(512, 2048), nn.ReLU(), nn.Linear(2048, 512)).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    queue = AsyncCallsQueue(persistent=args.persistent)
    for step in range(args.steps):
        model(torch.randn(8, 512, device=dev)).sum().backward()
        opt.step()
        opt.zero_grad()
        # every rank contributes its own keys plus a replicated one (deduplicated by the planner)
        state = {f"rank{rank}": model.state_dict(), "step": torch.tensor(step)}
        ckpt_dir = Path(args.ckpt_dir) / f"step{step}"
        if rank == 0 and ckpt_dir.exists():
            shutil.rmtree(ckpt_dir)
        dist.barrier()

        writer = FileSystemWriterAsync(ckpt_dir, thread_count=2)
        ret = save_state_dict_async_plan(state, writer, None, 0, planner=DefaultSavePlanner(), enable_cache=args.cache)
        save_fn, preload_fn, save_args = writer.get_save_function_and_args()
        finalize = [lambda ret=ret: save_state_dict_async_finalize(*ret)]
        queue.schedule_async_request(AsyncRequest(save_fn, save_args, finalize, preload_fn=preload_fn))
        # ... the next training step runs here while the writer process produces the files ...
        queue.maybe_finalize_async_calls(blocking=True)

        expect = {k: v.clone() for k, v in model.state_dict().items()}
        loaded = {f"rank{rank}": {k: torch.zeros_like(v) for k, v in expect.items()}, "step": torch.tensor(-1)}
        dcp.load(loaded, storage_reader=FileSystemReader(ckpt_dir))
        assert all(torch.equal(loaded[f"rank{rank}"][k], v) for k, v in expect.items()) and int(loaded["step"]) == step
        if rank == 0:
            print(f"step {step}: saved and verified {ckpt_dir}", flush=True)
    queue.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
