set -u
mkdir -p gpurun_out/v1
make -C nvidia-resiliency-ext_b200/csrc -j8 > gpurun_out/v1/build.log 2>&1 || { tail -20 gpurun_out/v1/build.log; exit 1; }
bash tools/gpu_round.sh tests
cp gpurun_out/pytest_gpu.log gpurun_out/v1/pytest_gpu_default.log
NVRX_B200_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_zcrc.py tests/test_gpu_zzero_copy.py tests/test_gpu_zdcp.py -m gpu -q --timeout=600 > gpurun_out/v1/pytest_gated.log 2>&1
tail -30 gpurun_out/v1/pytest_gated.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/v1/$name.json 2> gpurun_out/v1/$name.err; tail -4 gpurun_out/v1/$name.err; cat gpurun_out/v1/$name.json; }
run bench_ref --impl reference --steps 5 --warmup 1
run bench_c2 --steps 5
AB="--steps 3 --load-reps 1 --no-cpu-baseline --no-ceiling"
NVRX_B200_ZIP_CRC=1 run bench_copy_crc $AB
NVRX_B200_WRITE_FALLOCATE=1 NVRX_B200_RESTORE_PREAD=1 run bench_pread_falloc $AB
NVRX_B200_ZERO_COPY=1 run bench_zc $AB
NVRX_B200_ZERO_COPY=1 NVRX_B200_GPU_CRC=1 run bench_zc_crc $AB
run bench_c3 --config c3 --steps 5 --no-cpu-baseline --no-ceiling
nvidia-smi topo -m > gpurun_out/v1/topo.txt 2>&1; lscpu | head -30 > gpurun_out/v1/lscpu.txt; numactl -H >> gpurun_out/v1/lscpu.txt 2>&1
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled >> gpurun_out/v1/lscpu.txt 2>&1; mount | grep shm >> gpurun_out/v1/lscpu.txt
