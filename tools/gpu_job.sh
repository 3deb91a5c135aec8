# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo smoke rc=$?; tail -4 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_d.err; echo bench rc=$?
tail -n 5 gpurun_out/bench_d.err
timeout 900 python bench.py --games 32768 --pool-fraction 0.45 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_gomoku32768_pool045.json 2> gpurun_out/bench_p.err; echo bench rc=$?
tail -n 5 gpurun_out/bench_p.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/r02_launches_gomoku16384.csv python bench.py --steps 3 --warmup 3 --presearch 8 --no-cpu-baseline --no-extras > gpurun_out/ncu_g.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:res_block_kernel -s 21 -c 1 -o gpurun_out/prof_block_r02 -f python bench.py --steps 3 --warmup 3 --presearch 8 --no-cpu-baseline --no-extras > gpurun_out/ncu_b.log 2>&1; echo ncu rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read())
print('default', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e'], v['positions_per_s'], v['net_frac_of_sustained_peak'], v['roofline']['frac'] if v['roofline'] else None)
print('cpu', d.get('cpu_baseline'))
p=json.loads(open('gpurun_out/r02_bench_gomoku32768_pool045.json').read())
print('pool', p['value'], p['ms_per_step'], p['positions_per_s'], p['hbm_bytes'], p['slot_pool'])
"
python tools/launch_summary.py gpurun_out/r02_launches_gomoku16384.csv
