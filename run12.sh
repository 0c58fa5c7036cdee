mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "host_api or golden_batched" > gpurun_out/pytest12.log 2>&1; echo "rc=$?" >> gpurun_out/pytest12.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --batch 65536 ) > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
( time timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --batch 65536 --no-cpu ) > gpurun_out/bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n1.log
