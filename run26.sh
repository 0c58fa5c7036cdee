mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu26.log
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p26.log 2>&1
timeout 300 python bench.py --n 512 --batch 8192 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_512c.log 2>&1
