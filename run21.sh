mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu21.log
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p21.log 2>&1
PM_BENCH_EXTRA_FLAGS=0x2000 timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p21np.log 2>&1
