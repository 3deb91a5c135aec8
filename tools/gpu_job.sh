# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py tests/test_full_size_gpu.py -x -q > gpurun_out/r02_pytest_net.log 2>&1; echo pytest-net rc=$?
tail -15 gpurun_out/r02_pytest_net.log
timeout 600 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/r02_bench_connect4_fused.json 2> gpurun_out/bench_c.err; echo bench rc=$?
timeout 600 python bench.py --config tictactoe --no-cpu-baseline > gpurun_out/r02_bench_tictactoe_fused.json 2> gpurun_out/bench_t.err; echo bench rc=$?
tail -3 gpurun_out/bench_c.err gpurun_out/bench_t.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_connect4_fused.csv python bench.py --config connect4 --steps 3 --warmup 3 --presearch 4 --no-cpu-baseline --no-extras > gpurun_out/ncu_c4.log 2>&1; echo ncu rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_tictactoe.csv python bench.py --config tictactoe --steps 3 --warmup 3 --presearch 4 --no-cpu-baseline --no-extras > gpurun_out/ncu_ttt.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 8 -c 1 -o gpurun_out/prof_mlp -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_mlp.log 2>&1; echo ncu rc=$?
cut -c1-200 gpurun_out/r02_bench_connect4_fused.json; echo; cut -c1-200 gpurun_out/r02_bench_tictactoe_fused.json; echo
python tools/launch_summary.py gpurun_out/r02_launches_connect4_fused.csv
python tools/launch_summary.py gpurun_out/r02_launches_tictactoe.csv
