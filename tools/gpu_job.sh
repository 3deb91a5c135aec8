# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 300 python tools/ab_compare.py libgaz_ab_noproj.so gomoku 700 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_proj" -s 8 -c 3 --csv --log-file gpurun_out/r02_launches_stemproj.csv python tools/quick_net_bench.py gomoku 16384 > gpurun_out/ncu_q.log 2>&1; echo ncu rc=$?
grep '^"' gpurun_out/r02_launches_stemproj.csv | awk -F'","' '{print substr($5,1,60), $(NF)}'
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02_bench_gomoku_v4.json 2> gpurun_out/bench_v4.err; echo bench rc=$?
tail -n 3 gpurun_out/bench_v4.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_gomoku_v4.json') if l.startswith('{')][-1])
print('gomoku', d['value'], d['ms_per_step'], d['e2e']['value'], d['positions_per_s'], d['roofline']['frac'], d['clocks'])
"
