# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py --config connect4 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_connect4_v7.json 2> gpurun_out/bench_c7.err; echo bench rc=$?
timeout 900 python bench.py --config tictactoe --no-extras --no-cpu-baseline > gpurun_out/r02_bench_tictactoe_v7.json 2> gpurun_out/bench_t7.err; echo bench rc=$?
timeout 900 python bench.py --config gumbel --no-extras --no-cpu-baseline > gpurun_out/r02_bench_gumbel_v7.json 2> gpurun_out/bench_g7.err; echo bench rc=$?
tail -n 3 gpurun_out/bench_c7.err gpurun_out/bench_t7.err gpurun_out/bench_g7.err
python -c "
import json
for f in ('r02_bench_connect4_v7','r02_bench_tictactoe_v7','r02_bench_gumbel_v7'):
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline'] and d['roofline']['frac'], d['net_tflops'], d['clocks'], d['gpu_launches'])
"
