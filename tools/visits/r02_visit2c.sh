# 1 GPU, short: re-run of the fixed tests, where does the reference's DCP writer suite hang (faulthandler), stall breakdown
set -u
O=gpurun_out/v2c; mkdir -p $O
make -C nvidia-resiliency-ext_b200/csrc -j8 > $O/build.log 2>&1 || { tail -20 $O/build.log; exit 1; }
NVRX_B200_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_zzero_copy.py tests/test_gpu_zcrc.py tests/test_gpu_zdcp.py tests/test_gpu_api.py -m gpu -q --timeout=300 > $O/pytest_refix2.log 2>&1
tail -6 $O/pytest_refix2.log | cut -c1-300
cd tests/golden/ref_tests
for T in test_async_is_equivalent_to_sync test_invalid_async_setup test_errors_are_reported test_cached_metadata test_cached_data_structure test_cpu_shm_for_gpu_tensors test_async_cp_with_multiple_queue_and_abort; do
  PYTHONPATH=$GRAFT_REPO_ROOT/nvidia-resiliency-ext_b200 timeout 150 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29655 \
    -m pytest -q -x -p no:cacheprovider --confcutdir . --rootdir . -o faulthandler_timeout=60 tests/checkpointing/unit/test_async_writer.py -k "$T" > $GRAFT_REPO_ROOT/$O/refdcp_$T.log 2>&1
  echo "$T exit=$?"; tail -3 $GRAFT_REPO_ROOT/$O/refdcp_$T.log | cut -c1-200
done
cd $GRAFT_REPO_ROOT
NVRX_B200_TRACE=1 timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-restore --no-ceiling --load-reps 1 > $O/bench_trace.json 2> $O/bench_trace.err; tail -2 $O/bench_trace.err; cat $O/bench_trace.json
