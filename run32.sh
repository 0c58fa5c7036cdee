mkdir -p gpurun_out
PM_LIB_VARIANT=256x3 timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p32.log 2>&1
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p32b.log 2>&1
