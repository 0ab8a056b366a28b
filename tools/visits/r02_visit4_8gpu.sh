# 8 GPUs (charged 8x): every minute counts.  Order = value: green multi-GPU log, C4/C5 at 16 GB/rank (both arms), fused
# all-to-all A/B, host ceiling, N=8 bench line.  Every step has its own timeout; outputs under gpurun_out/v4.
# The library is NOT rebuilt here (a minute of 8 GPUs): the in-tree .so was built from this tree; the symbol check fails fast.
set -u
O=gpurun_out/v4; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'nvidia-resiliency-ext_b200')
from nvidia_resiliency_ext.checkpointing.b200 import _cabi
l=_cabi.lib(); assert all(hasattr(l,n) for n in _cabi.EXPORTED_SYMBOLS); print('library ok, abi', l.nvrx_abi_version())" || exit 1
{ echo "memory.max: $(cat /sys/fs/cgroup/memory.max 2>/dev/null)"; nproc; nvidia-smi topo -m; } > $O/box.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NVRX_B200_TEST_UNVALIDATED=1 timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=10 --timeout=240 > $O/pytest_multi_8gpu.log 2>&1
tail -8 $O/pytest_multi_8gpu.log | cut -c1-250
run() { name=$1; port=$2; to=$3; shift 3; timeout $to $T --master-port $port bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; tail -3 $O/$name.err | cut -c1-250; cat $O/$name.json; rm -rf /dev/shm/nvrx_b200_* /dev/shm/nvrx_* 2>/dev/null; }
run c4_engine 29803 420 --config c4 --c4-iters 1 --c4-warm 1 --layouts pairs,striped,full
run c4_reference 29804 240 --config c4 --impl reference --c4-iters 1 --c4-warm 1 --layouts pairs
timeout 150 $T --master-port 29801 tools/bench_replicate.py --kernel-only --scale 1.0 --iters 3 > $O/replk_interleave.json 2> $O/replk_interleave.err; cat $O/replk_interleave.json
NVRX_B200_SHARD_INTERLEAVE=0 timeout 150 $T --master-port 29802 tools/bench_replicate.py --kernel-only --scale 1.0 --iters 3 > $O/replk_sequential.json 2> $O/replk_sequential.err; cat $O/replk_sequential.json
timeout 150 $T --master-port 29805 tools/d2h_ceiling.py --gb 8 --reps 2 > $O/d2h_ceiling.jsonl 2> $O/d2h_ceiling.err; cat $O/d2h_ceiling.jsonl
run bench_n8 29806 240 --steps 3 --load-reps 1 --no-restore
ls -la $O
