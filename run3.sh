mkdir -p gpurun_out
for v in "" 256x2 256x3; do
  PM_LIB_VARIANT=$v timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p_$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_p_$v.log
done
