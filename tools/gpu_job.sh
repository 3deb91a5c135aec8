# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/bench_df.err; echo bench rc=$?
echo "default bench wall seconds: $(( $(date +%s) - t0 ))"
tail -n 3 gpurun_out/bench_df.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default_final.json') if l.startswith('{')][-1])
print('default', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['roofline']['traffic'], d['net_tflops'], d['clocks'], d['gpu_launches'])
for k,v in d.get('configs',{}).items(): print(k, json.dumps({kk:vv for kk,vv in v.items() if kk not in ('workload','whole_move','clocks')})[:600])
print('cpu', d.get('cpu_baseline'))
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
