mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
