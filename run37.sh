mkdir -p gpurun_out
PM_LIB_VARIANT=512x2 timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_p37.log 2>&1
