# scratch driver for one gpurun call: full round-end verification (what the driver runs)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); r=d['roofline']; print('ours', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['clocks']['sm_mhz'], round(d['cpu_baseline']['value']), d['gpu_launches'])"
