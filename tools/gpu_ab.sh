# scratch driver for one gpurun call: whole-game pool sizing - default bench, other configs, Gomoku PUCT generation
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/caps_gomoku.json 2> gpurun_out/caps_gomoku.err; echo rc=$?
for c in connect4 gumbel tictactoe; do timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/caps_$c.json 2> gpurun_out/caps_$c.err; echo rc=$?; done
for f in gomoku connect4 gumbel tictactoe; do tail -c 300 gpurun_out/caps_$f.err; python -c "
import json; d=json.loads(open('gpurun_out/caps_$f.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['frac'], d['clocks']['sm_mhz'], d['hbm_bytes'])"; done
timeout 900 python tools/selfplay_generation.py gomoku 256 /tmp/g3 > gpurun_out/gen_gomoku.json 2> gpurun_out/gen_gomoku.err; echo rc=$?; tail -c 700 gpurun_out/gen_gomoku.json; tail -3 gpurun_out/gen_gomoku.err
