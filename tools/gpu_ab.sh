# scratch driver for one gpurun call: dual head convolution (Connect4)
timeout 120 python tools/check_head_dual.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py -q -x 2>&1 | tail -3
for v in 1 0 1 0; do GAZ_HEAD_DUAL=$v timeout 300 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/c4_dual$v.json 2> gpurun_out/c4_dual$v.err; python -c "
import json; d=json.loads(open('gpurun_out/c4_dual$v.json').read().strip().splitlines()[-1]); print('dual=$v', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'])"; done
