timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo rc=$?
for c in connect4 gumbel tictactoe; do timeout 400 python bench.py --config $c > gpurun_out/bench_${c}4.json 2> gpurun_out/bench_${c}4.err; echo rc=$?; done
for f in final4 connect44 gumbel4 tictactoe4; do python -c "
import json; d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['clocks']['sm_mhz'], round(d['cpu_baseline']['value']), d['gpu_launches'])"; done
