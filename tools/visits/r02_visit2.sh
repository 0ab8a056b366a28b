set -u
O=gpurun_out/v2; mkdir -p $O
make -C nvidia-resiliency-ext_b200/csrc -j8 > $O/build.log 2>&1 || { tail -20 $O/build.log; exit 1; }
NVRX_B200_TEST_UNVALIDATED=1 timeout 1500 python -m pytest tests -m gpu -q --durations=12 --timeout=900 > $O/pytest_gpu_all.log 2>&1
tail -40 $O/pytest_gpu_all.log | cut -c1-400
timeout 600 python tools/restore_breakdown.py > $O/restore_breakdown.json 2> $O/restore_breakdown.err; tail -3 $O/restore_breakdown.err; cat $O/restore_breakdown.json
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/$name.json 2> $O/$name.err; tail -4 $O/$name.err; cat $O/$name.json; }
run bench_ref --impl reference --steps 5 --warmup 1
run bench_c2 --steps 10
run bench_c3 --config c3 --steps 5 --no-cpu-baseline --no-ceiling
NVRX_B200_ZERO_COPY=0 run bench_c2_copy --steps 3 --load-reps 1 --no-cpu-baseline --no-ceiling
