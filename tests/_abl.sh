timeout 600 python -m pytest tests/test_net_gpu.py tests/test_facade_gpu.py -q -s 2>&1 | grep -E "connect4 s|tictactoe|passed|failed" | sort | uniq -c
timeout 400 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/bench_c4_v3.json 2> gpurun_out/bench_c4_v3.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_v3.json').read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['gpu_launches'])"
