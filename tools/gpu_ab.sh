# scratch driver for one gpurun call (A/B of bench variants on one box)
for c in gomoku connect4 gumbel tictactoe; do timeout 500 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_${c}_noise.json 2> gpurun_out/bench_${c}_noise.err; echo rc=$?; done
timeout 500 python bench.py --config gomoku --no-noise --no-cpu-baseline > gpurun_out/bench_gomoku_nonoise.json 2> gpurun_out/bench_gomoku_nonoise.err; echo rc=$?
for f in gomoku_noise gomoku_nonoise connect4_noise gumbel_noise tictactoe_noise; do python -c "
import json; d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['clocks']['sm_mhz'], d['sims_per_eval'], d['gpu_launches'])"; done
tail -3 gpurun_out/bench_*_noise.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 90 --csv --log-file gpurun_out/launches_noise.csv python bench.py --steps 3 --warmup 3 --presearch 8 --no-cpu-baseline > gpurun_out/ncu_noise.log 2>&1; echo ncu rc=$?
