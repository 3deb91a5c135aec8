# scratch driver for one gpurun call: zero-skipping CUDA-core stem (Connect4 / TicTacToe)
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_net_golden_gpu.py -q -x 2>&1 | tail -3
for v in 1 0 1 0; do GAZ_STEM_SKIP0=$v timeout 300 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/c4_skip$v.json 2> gpurun_out/c4_skip$v.err; python -c "
import json; d=json.loads(open('gpurun_out/c4_skip$v.json').read().strip().splitlines()[-1]); print('skip0=$v', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'])"; done
for v in 1 0; do GAZ_STEM_SKIP0=$v timeout 300 python bench.py --config tictactoe --no-cpu-baseline > gpurun_out/ttt_skip$v.json 2> gpurun_out/ttt_skip$v.err; python -c "
import json; d=json.loads(open('gpurun_out/ttt_skip$v.json').read().strip().splitlines()[-1]); print('ttt skip0=$v', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'])"; done
