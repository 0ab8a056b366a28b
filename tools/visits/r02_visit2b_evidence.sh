# 1 GPU: refresh the evidence set -- ncu launch list + full capture of the current walkers, compute-sanitizer memcheck / racecheck
set -u
O=gpurun_out/v2b; mkdir -p $O
make -C nvidia-resiliency-ext_b200/csrc -j8 > $O/build.log 2>&1 || { tail -20 $O/build.log; exit 1; }
NVRX_B200_TEST_UNVALIDATED=1 timeout 1500 python -m pytest tests/test_gpu_api.py tests/test_gpu_zzero_copy.py tests/test_gpu_zcrc.py tests/test_gpu_zzz_reference_suite.py -m gpu -q --timeout=900 > $O/pytest_refix.log 2>&1
tail -12 $O/pytest_refix.log | cut -c1-300
timeout 600 python tools/restore_breakdown.py > $O/restore_breakdown.json 2> $O/restore_breakdown.err; tail -3 $O/restore_breakdown.err | cut -c1-300; cat $O/restore_breakdown.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk_ -c 80 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify --no-restore --no-training-loop --no-ceiling > $O/bench_under_ncu.log 2>&1
grep -c walk_ $O/launches.csv
NCU_REPS=1 NCU_CRC=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_ -c 6 -f -o $O/prof_walk \
    python tools/ncu_target.py > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
K="ragged or roundtrip_into or narrow or degenerate or fill_from_fd"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > $O/sanitizer_memcheck.log 2>&1
echo "memcheck exit code: $?" >> $O/sanitizer_memcheck.log; tail -6 $O/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > $O/sanitizer_racecheck.log 2>&1
echo "racecheck exit code: $?" >> $O/sanitizer_racecheck.log; tail -6 $O/sanitizer_racecheck.log
ls -la $O
