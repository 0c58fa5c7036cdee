set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --batch 16384 --steps 3 --warmup 3 > gpurun_out/bench_16k.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_16k.log
timeout 300 python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/b_small.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/b_small2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pm_analyse -s 3 -c 1 -o gpurun_out/prof_r1a python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_16k.log
