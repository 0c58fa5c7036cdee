mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
for v in "" 256x2; do
  PM_LIB_VARIANT=$v timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p8_$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_p8_$v.log
done
