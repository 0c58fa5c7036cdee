mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu34.log
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p34.log 2>&1
