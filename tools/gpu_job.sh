# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
AB_ORACLE=1 timeout 600 python tools/ab_compare.py libgaz_ab_noheadtc.so connect4 600 2>&1 | tail -7
AB_ORACLE=1 timeout 600 python tools/ab_compare.py libgaz_ab_noheadtc.so connect4 5 2>&1 | tail -7
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_net.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r02_pytest_net.log
timeout 300 python tools/quick_net_bench.py connect4 4096 2>&1 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/r02_launches_net_connect4_v6.csv python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_q.log 2>&1; echo ncu rc=$?
grep '^"' gpurun_out/r02_launches_net_connect4_v6.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | tail -12
