set -x
mkdir -p gpurun_out
timeout 300 python tools/ab_compare.py libgaz_ab_nostreams.so gomoku 300 2>&1 | tail -5
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py tests/test_net_fusion_gpu.py tests/test_eval_cache.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/quick_net_bench.py gomoku 16384 2>&1 | tail -4
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02_bench_gomoku_v9.json 2> gpurun_out/bench_v9.err; echo bench rc=$?
tail -3 gpurun_out/bench_v9.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_gomoku_v9.json') if l.startswith('{')][-1])
print('gomoku', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
"
