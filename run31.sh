mkdir -p gpurun_out
timeout 300 python tests/sanitizer_case.py > gpurun_out/san_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tests/sanitizer_case.py > gpurun_out/san_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck.log
