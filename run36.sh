mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_full36.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full36.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_n2_36.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_36.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke36.log 2>&1; echo "rc=$?" >> gpurun_out/smoke36.log
