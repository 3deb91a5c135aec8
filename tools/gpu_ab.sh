# scratch driver for one gpurun call: whole-move timing (positions/s measured directly)
timeout 300 python bench.py --config tictactoe --moves 4 > gpurun_out/moves_tictactoe.json 2> gpurun_out/moves_tictactoe.err; echo rc=$?
timeout 400 python bench.py --config connect4 --moves 6 > gpurun_out/moves_connect4.json 2> gpurun_out/moves_connect4.err; echo rc=$?
timeout 400 python bench.py --config gumbel --moves 6 > gpurun_out/moves_gumbel.json 2> gpurun_out/moves_gumbel.err; echo rc=$?
timeout 600 python bench.py --config gomoku --moves 2 > gpurun_out/moves_gomoku.json 2> gpurun_out/moves_gomoku.err; echo rc=$?
for f in tictactoe connect4 gumbel gomoku; do tail -c 400 gpurun_out/moves_$f.err; python -c "
import json; d=json.loads(open('gpurun_out/moves_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), d['steps'], round(d['positions_per_s'],1), d['moves_timed'], d['games_still_running'], round(d['sims_per_eval'],3), d['clocks']['sm_mhz'])"; done
