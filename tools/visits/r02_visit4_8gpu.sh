# 8 GPUs (charged 8x): every minute counts.  Order = value: green multi-GPU log, fused all-to-all A/B, C4/C5 at 16 GB/rank (both
# arms), host ceiling, N=8 bench lines.  Every step has its own timeout; outputs under gpurun_out/v4.
set -u
O=gpurun_out/v4; mkdir -p $O
make -C nvidia-resiliency-ext_b200/csrc -j16 > $O/build.log 2>&1 || { tail -20 $O/build.log; exit 1; }
{ echo "memory.max: $(cat /sys/fs/cgroup/memory.max 2>/dev/null)"; nproc; nvidia-smi topo -m; } > $O/box.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NVRX_B200_TEST_UNVALIDATED=1 timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=10 --timeout=300 > $O/pytest_multi_8gpu.log 2>&1
tail -8 $O/pytest_multi_8gpu.log | cut -c1-250
timeout 200 $T --master-port 29801 tools/bench_replicate.py --kernel-only --scale 1.0 --iters 3 > $O/replk_interleave.json 2> $O/replk_interleave.err; cat $O/replk_interleave.json
NVRX_B200_SHARD_INTERLEAVE=0 timeout 200 $T --master-port 29802 tools/bench_replicate.py --kernel-only --scale 1.0 --iters 3 > $O/replk_sequential.json 2> $O/replk_sequential.err; cat $O/replk_sequential.json
run() { name=$1; port=$2; to=$3; shift 3; timeout $to $T --master-port $port bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; tail -3 $O/$name.err | cut -c1-250; cat $O/$name.json; rm -rf /dev/shm/nvrx_b200_* /dev/shm/nvrx_* 2>/dev/null; }
run c4_engine 29803 420 --config c4 --c4-iters 1 --c4-warm 1 --layouts pairs,striped,full
NVRX_B200_EXCHANGE=p2p run c4_engine_striped_fused 29808 240 --config c4 --c4-iters 1 --c4-warm 1 --layouts striped
run c4_reference 29804 300 --config c4 --impl reference --c4-iters 1 --c4-warm 1 --layouts pairs,full
timeout 200 $T --master-port 29805 tools/d2h_ceiling.py --gb 8 --reps 2 > $O/d2h_ceiling.jsonl 2> $O/d2h_ceiling.err; cat $O/d2h_ceiling.jsonl
run bench_n8 29806 300 --steps 3 --load-reps 1
run bench_n8_ref 29807 300 --impl reference --steps 3 --warmup 1 --load-reps 1
ls -la $O
