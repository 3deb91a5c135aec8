# scratch driver for one gpurun call: Gomoku generations (PUCT 256 games, Gumbel 1024 games) through run_self_play
timeout 500 python tools/selfplay_generation.py gumbel 1024 /tmp/g2 > gpurun_out/gen_gumbel.json 2> gpurun_out/gen_gumbel.err; echo rc=$?; tail -c 700 gpurun_out/gen_gumbel.json; tail -3 gpurun_out/gen_gumbel.err
timeout 700 python tools/selfplay_generation.py gomoku 256 /tmp/g3 > gpurun_out/gen_gomoku.json 2> gpurun_out/gen_gomoku.err; echo rc=$?; tail -c 700 gpurun_out/gen_gomoku.json; tail -3 gpurun_out/gen_gomoku.err
