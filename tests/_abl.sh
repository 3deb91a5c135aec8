timeout 300 python tests/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python tests/sanitize_small.py > gpurun_out/san_run.log 2>&1
echo rc=$?; tail -3 gpurun_out/san_plain.log; tail -5 gpurun_out/san_run.log; tail -5 gpurun_out/memcheck.log
