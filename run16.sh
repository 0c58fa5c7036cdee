mkdir -p gpurun_out
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p16.log 2>&1; echo "rc=$?" >> gpurun_out/bench_p16.log
