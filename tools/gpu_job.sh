set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --config connect4 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_connect4_v10.json 2> gpurun_out/bench_c10.err; echo bench rc=$?
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_connect4_v10.json') if l.startswith('{')][-1])
print('connect4', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['net_tflops'], d['clocks'], d['gpu_launches'])
"
