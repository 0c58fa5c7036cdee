mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu23.log
timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p23.log 2>&1
