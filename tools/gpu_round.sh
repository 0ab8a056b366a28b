#!/bin/bash
# One GPU-box visit: tests, bench (both arms), ncu launch list + full capture.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
{ echo "memory.max: $(cat /sys/fs/cgroup/memory.max 2>/dev/null)"; echo "memory.current: $(cat /sys/fs/cgroup/memory.current 2>/dev/null)"; df -h /dev/shm | tail -1; } > gpurun_out/memlimits.txt 2>&1
WHAT=${1:-all}

# Memory guard: the box's cgroup (memory.max, 200 GiB on a 1-GPU box) also counts tmpfs / POSIX-shm pages, i.e. the
# pinned snapshot slots and the checkpoint files in /dev/shm.  Exceeding it kills the whole box.  Every workload below
# runs in its own session; a watcher kills that session (only it) and removes our shm files above 80 % of the limit.
MEM_MAX=$(cat /sys/fs/cgroup/memory.max 2>/dev/null || echo max)
[[ $MEM_MAX == max ]] && MEM_MAX=$(( $(grep MemTotal /proc/meminfo | awk '{print $2}') * 1024 ))
MEM_CAP=$(( MEM_MAX / 100 * ${MEM_GUARD_PCT:-80} ))
guarded() {
  setsid "$@" &
  local pid=$!
  ( while kill -0 $pid 2>/dev/null; do
      cur=$(cat /sys/fs/cgroup/memory.current 2>/dev/null || echo 0)
      if (( cur > MEM_CAP )); then
        echo "MEMORY GUARD: $cur > $MEM_CAP, killing session $pid" >> gpurun_out/memguard.txt
        kill -TERM -- -$pid 2>/dev/null; sleep 2; kill -KILL -- -$pid 2>/dev/null
        rm -f /dev/shm/nvrx_* 2>/dev/null; rm -rf /dev/shm/nvrx_b200_* 2>/dev/null
        break
      fi
      sleep 0.5
    done ) &
  wait $pid
}
if [[ $WHAT == all || $WHAT == tests ]]; then
  guarded timeout 1500 python -m pytest tests -m gpu -q --durations=15 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
  tail -25 gpurun_out/pytest_gpu.log
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  guarded timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
  guarded timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [[ $WHAT == zerocopy ]]; then
  # zero-copy persistence (default since round 2), GPU checksums (opt-in), DCP async writer: tests, then the bench in each writer mode
  guarded timeout 900 python -m pytest tests/test_gpu_zcrc.py tests/test_gpu_zzero_copy.py tests/test_gpu_zdcp.py -m gpu -q --timeout=600 > gpurun_out/pytest_zerocopy.log 2>&1
  tail -25 gpurun_out/pytest_zerocopy.log
  guarded timeout 600 python tools/crc_bench.py --gb 2 > gpurun_out/crc_bench.jsonl 2> gpurun_out/crc_bench.err
  guarded timeout 600 python tools/crc_bench.py --gb 16 >> gpurun_out/crc_bench.jsonl 2>> gpurun_out/crc_bench.err; cat gpurun_out/crc_bench.jsonl
  NVRX_B200_ZERO_COPY=0 guarded timeout 900 python bench.py > gpurun_out/bench_copy.json 2> gpurun_out/bench_copy.err; cat gpurun_out/bench_copy.json
  NVRX_B200_ZERO_COPY=0 NVRX_B200_NO_FALLOCATE=1 guarded timeout 900 python bench.py > gpurun_out/bench_copy_nofallocate.json 2> gpurun_out/bench_copy_nofallocate.err; cat gpurun_out/bench_copy_nofallocate.json
  NVRX_B200_RESTORE_PREAD=0 guarded timeout 900 python bench.py > gpurun_out/bench_mmap_restore.json 2> gpurun_out/bench_mmap_restore.err; cat gpurun_out/bench_mmap_restore.json
  NVRX_B200_GPU_CRC=1 guarded timeout 900 python bench.py > gpurun_out/bench_zerocopy.json 2> gpurun_out/bench_zerocopy.err; tail -3 gpurun_out/bench_zerocopy.err; cat gpurun_out/bench_zerocopy.json
fi
if [[ $WHAT == multizc ]]; then
  # 2+ GPUs: the whole multi-GPU suite including the gated replicated zero-copy test
  guarded timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=10 --timeout=900 > gpurun_out/pytest_multi_zc.log 2>&1
  tail -30 gpurun_out/pytest_multi_zc.log
fi
if [[ $WHAT == multi ]]; then
  guarded timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=10 --timeout=900 > gpurun_out/pytest_multi.log 2>&1
  tail -30 gpurun_out/pytest_multi.log
fi
if [[ $WHAT == repl ]]; then
  NG=${2:-2}; SCALE=${3:-0.25}; LAYOUTS=${4:-"full sharded"}
  for LAYOUT in $LAYOUTS; do for MODE in p2p nccl; do
    guarded timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2961$NG \
        tools/bench_replicate.py --scale $SCALE --mode $MODE --layout $LAYOUT --iters 3 > gpurun_out/repl_${NG}_${LAYOUT}_${MODE}.json 2> gpurun_out/repl_${NG}_${LAYOUT}_${MODE}.err
    tail -2 gpurun_out/repl_${NG}_${LAYOUT}_${MODE}.err; cat gpurun_out/repl_${NG}_${LAYOUT}_${MODE}.json
  done; done
fi
if [[ $WHAT == replk ]]; then
  NG=${2:-2}; SCALE=${3:-1.0}
  guarded timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2963$NG \
      tools/bench_replicate.py --kernel-only --scale $SCALE --iters 5 > gpurun_out/replk_${NG}.json 2> gpurun_out/replk_${NG}.err
  tail -3 gpurun_out/replk_${NG}.err; cat gpurun_out/replk_${NG}.json
fi
if [[ $WHAT == scale ]]; then
  NG=${2:-2}
  guarded timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2962$NG \
      bench.py --gpus $NG > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
  tail -2 gpurun_out/bench_n$NG.err; cat gpurun_out/bench_n$NG.json
fi
if [[ $WHAT == sanitize ]]; then
  # memcheck over the ragged (funnel / head / tail) paths and both walkers: no out-of-bounds access on odd alignments
  guarded timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -q -x -k "ragged or roundtrip_into or narrow or degenerate" > gpurun_out/sanitizer_memcheck.log 2>&1
  echo "memcheck exit code: $?" >> gpurun_out/sanitizer_memcheck.log
  tail -8 gpurun_out/sanitizer_memcheck.log
fi
if [[ $WHAT == ncu_list ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk_ -c 40 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify --no-restore --no-training-loop --no-ceiling > gpurun_out/bench_under_ncu.log 2>&1
  grep -c walk_ gpurun_out/launches.csv
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk_ -c 40 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify --no-restore --no-training-loop --no-ceiling > gpurun_out/bench_under_ncu.log 2>&1
  grep -c walk_ gpurun_out/launches.csv
  NCU_REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_ -c 6 -f -o gpurun_out/prof_walk \
      python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
fi
if [[ $WHAT == ncu_crc ]]; then
  NCU_REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:crc_chunks -c 2 -f -o gpurun_out/prof_crc \
      python tools/ncu_target.py > gpurun_out/ncu_crc.log 2>&1
  tail -3 gpurun_out/ncu_crc.log; ls -la gpurun_out/
fi
