mkdir -p gpurun_out
for v in 320x2 384x2; do
PM_LIB_VARIANT=$v timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_p35_$v.log 2>&1
done
