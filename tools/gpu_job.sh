set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/r02_launches_net_connect4_v5.csv python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_q.log 2>&1; echo ncu rc=$?
grep '^"' gpurun_out/r02_launches_net_connect4_v5.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | tail -12
timeout 900 python bench.py --config connect4 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_connect4_v6.json 2> gpurun_out/bench_c6.err; echo bench rc=$?
python -c "
import json
for f in ('r02_bench_connect4_v6',):
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['net_tflops'], d['clocks'], d['gpu_launches'])
"
