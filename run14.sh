mkdir -p gpurun_out
timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p14.log 2>&1
PM_LIB_VARIANT=512x1 timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p14b.log 2>&1
