mkdir -p gpurun_out
timeout 300 python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/b_small20.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pm_analyse -s 3 -c 1 -o gpurun_out/prof_r1e python bench.py --batch 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_e.log 2>&1
