mkdir -p gpurun_out
for sm in 98304 81920 65536; do
PM_SMEM_BYTES=$sm timeout 300 python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu --phases > gpurun_out/bench_p19_$sm.log 2>&1; echo "rc=$?" >> gpurun_out/bench_p19_$sm.log
done
