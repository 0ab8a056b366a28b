# 2 GPUs (charged 2x): multi-GPU suite incl. gated tests, reference unit tests at world 2, C4 script check at reduced scale, N=2 bench
set -u
O=gpurun_out/v3; mkdir -p $O
make -C nvidia-resiliency-ext_b200/csrc -j8 > $O/build.log 2>&1 || { tail -20 $O/build.log; exit 1; }
NVRX_B200_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=12 --timeout=600 > $O/pytest_multi_2gpu.log 2>&1
tail -30 $O/pytest_multi_2gpu.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_zzz_reference_suite.py -m gpu -q --timeout=700 -k "two_ranks and not dcp" > $O/pytest_refsuite_world2.log 2>&1
tail -6 $O/pytest_refsuite_world2.log | cut -c1-300
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; port=$2; shift 2; timeout 900 $T --master-port $port bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; tail -4 $O/$name.err | cut -c1-300; cat $O/$name.json; }
run c4_engine 29701 --config c4 --scale 0.25 --c4-iters 1 --c4-warm 1
NVRX_B200_ZERO_COPY_REPLICAS=1 run c4_engine_zcr 29702 --config c4 --scale 0.25 --c4-iters 2 --c4-warm 1 --layouts pairs
NVRX_B200_EXCHANGE=stream run c4_engine_stream 29703 --config c4 --scale 0.25 --c4-iters 1 --c4-warm 1 --layouts pairs
run c4_reference 29704 --config c4 --impl reference --scale 0.25 --c4-iters 1 --c4-warm 1
run bench_n2 29705 --steps 3 --load-reps 1
run bench_n2_ref 29706 --impl reference --steps 3 --warmup 1 --load-reps 1
