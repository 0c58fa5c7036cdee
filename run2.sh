set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
for v in "" 256x2 256x3 256x4 1024x1; do
  PM_LIB_VARIANT=$v timeout 300 python bench.py --batch 16384 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v_$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_v_$v.log
done
timeout 300 python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/b_small.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pm_analyse -s 3 -c 1 -o gpurun_out/prof_r1b python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_b.log 2>&1
