# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py tests/test_self_play_reference_golden.py tests/test_self_play_cpu.py tests/test_session_cache.py -x -q > gpurun_out/r02_pytest_net.log 2>&1; echo pytest-net rc=$?
tail -5 gpurun_out/r02_pytest_net.log
timeout 600 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/r02_bench_connect4_small.json 2> gpurun_out/bench_c.err; echo bench rc=$?
timeout 600 python bench.py --config tictactoe --no-cpu-baseline > gpurun_out/r02_bench_tictactoe_small.json 2> gpurun_out/bench_t.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_connect4_small.csv python bench.py --config connect4 --steps 3 --warmup 3 --presearch 4 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:headconv_wide -s 3 -c 1 -o gpurun_out/prof_headwide -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_hw.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 3 -c 1 -o gpurun_out/prof_stemmma -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_st.log 2>&1; echo ncu rc=$?
cut -c1-200 gpurun_out/r02_bench_connect4_small.json; echo; cut -c1-200 gpurun_out/r02_bench_tictactoe_small.json; echo
python tools/launch_summary.py gpurun_out/r02_launches_connect4_small.csv
