timeout 300 python -m pytest tests/test_net_gpu.py -q -s 2>&1 | grep -E "connect4 softmax|passed|failed" | sort | uniq -c
timeout 400 python bench.py --config connect4 --no-cpu-baseline > gpurun_out/bench_c4_f32v.json 2> gpurun_out/bench_c4_f32v.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_f32v.json').read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['clocks']['sm_mhz'])"
GAZ_HEAD_F32V=0 timeout 400 python bench.py --config connect4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic', round(d['value']), round(d['ms_per_step'],3))"
