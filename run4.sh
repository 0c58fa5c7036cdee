mkdir -p gpurun_out
timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_p4.log
