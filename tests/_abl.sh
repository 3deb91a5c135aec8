timeout 600 python -m pytest tests/test_net_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tests/quick_net_bench.py gomoku 16384 > gpurun_out/qnb.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_head.csv python tests/quick_net_bench.py gomoku 16384 > gpurun_out/ncu_head.log 2>&1; echo rc=$?
