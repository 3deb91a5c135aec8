timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_v3.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], {k:(v['avg_launch_ms'],v['launches']) for k,v in r['groups'].items()}, d['clocks'])"
L=grok_alpha_zero_b200/libgaz_b200.so
cp $L /tmp/head.so; cp tests/_emul/libgaz_prev.so $L
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_prev.json 2> gpurun_out/bench_prev.err; echo rc=$?
cp /tmp/head.so $L
python -c "
import json; d=json.loads(open('gpurun_out/bench_prev.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], {k:(v['avg_launch_ms'],v['launches']) for k,v in r['groups'].items()}, d['clocks'])"
