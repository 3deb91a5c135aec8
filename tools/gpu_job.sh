# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/net_undamped.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r02_smoke.log
timeout 600 python tools/net_tolerance_sweep.py > gpurun_out/r02_net_tolerance.jsonl 2> gpurun_out/sweep.err; echo sweep rc=$?
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_gomoku_base.json 2> gpurun_out/bench_g.err; echo bench rc=$?
timeout 600 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/r02_bench_connect4_base.json 2> gpurun_out/bench_c.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_connect4_base.csv python bench.py --config connect4 --steps 3 --warmup 3 --presearch 4 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; echo ncu rc=$?
cut -c1-400 gpurun_out/r02_bench_gomoku_base.json; echo; cut -c1-400 gpurun_out/r02_bench_connect4_base.json
