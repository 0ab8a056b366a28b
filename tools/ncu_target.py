"""Small launch sequence for ncu: the C2 state, then pack (TMA walker), pack (LDG walker), scatter, narrow pack, and the
checksum kernel in both table layouts (NCU_CRC=0 skips it).
Usage under ncu:  ncu --set full -k regex:walk_ -c 6 -o gpurun_out/prof python tools/ncu_target.py
                  ncu --set full -k regex:crc_chunks -c 2 -o gpurun_out/prof_crc python tools/ncu_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nvidia-resiliency-ext_b200"))
import torch  # noqa: E402

from bench import flatten, llama3_8b_shard_state  # noqa: E402
from nvidia_resiliency_ext.checkpointing.b200.engine import DeviceBuffer, Plan  # noqa: E402

scale = float(os.environ.get("NCU_SCALE", "1.0"))
sd, total = llama3_8b_shard_state(torch.device("cuda"), scale=scale)
tensors = flatten(sd)
ptrs = [t.data_ptr() for t in tensors]
nbytes = [t.numel() * 4 for t in tensors]
st = torch.cuda.current_stream().cuda_stream
plan = Plan(ptrs, nbytes, None, device=0)
stg = DeviceBuffer(plan.staging_bytes, 0)
reps = int(os.environ.get("NCU_REPS", "2"))
for variant in (2, 1):
    plan.set_variant(variant)
    for _ in range(reps):
        plan.pack(stg.ptr, st)
    for _ in range(reps):
        plan.scatter(stg.ptr, st)
torch.cuda.synchronize()
narrow = Plan(ptrs, nbytes, [1 if t.numel() > 1 else 0 for t in tensors], device=0)
for _ in range(reps):
    narrow.pack(stg.ptr, st)
for _ in range(reps):
    narrow.scatter(stg.ptr, st)
torch.cuda.synchronize()
if os.environ.get("NCU_CRC", "1") != "0":
    from nvidia_resiliency_ext.checkpointing.b200.engine import CrcPlan  # noqa: E402

    plan.set_variant(2)
    plan.pack(stg.ptr, st)
    crc = CrcPlan(plan.offsets, plan.packed_nbytes, 0)
    values = torch.zeros(crc.n_values + 2, dtype=torch.int32).pin_memory()
    for variant in ("private", "shared"):
        os.environ["NVRX_B200_CRC_VARIANT"] = variant
        for _ in range(reps):
            crc.run(stg.ptr, values.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
print("ncu target done", total)
