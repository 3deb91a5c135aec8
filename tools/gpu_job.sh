# driver of one gpurun call: the round-end verification on a fresh B200 box
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_job.sh'
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/smoke.log
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
tail -n 3 gpurun_out/bench_default.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('default', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['gpu_launches'])
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'))
print('cpu', d.get('cpu_baseline'))
"
