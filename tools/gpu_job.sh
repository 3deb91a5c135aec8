set -x
mkdir -p gpurun_out
AB_ORACLE=1 timeout 300 python tools/ab_compare.py libgaz_ab_nomlpmma.so gomoku 100 2>&1 | tail -6
AB_ORACLE=1 timeout 300 python tools/ab_compare.py libgaz_ab_nomlpmma.so connect4 600 2>&1 | tail -6
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py tests/test_net_fusion_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/quick_net_bench.py gomoku 16384 2>&1 | tail -4
timeout 300 python tools/quick_net_bench.py connect4 4096 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mlp_chain|dense_kernel" -s 16 -c 4 --csv --log-file gpurun_out/mlp.csv python tools/quick_net_bench.py gomoku 16384 > gpurun_out/ncu_q.log 2>&1; echo ncu rc=$?
grep '^"' gpurun_out/mlp.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | tail -5
