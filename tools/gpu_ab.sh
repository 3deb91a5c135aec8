# scratch driver for one gpurun call: tensor-core TicTacToe net + full generations through run_self_play
timeout 300 python -m pytest tests/test_net_gpu.py -q -s -k "tictactoe" 2>&1 | grep -E "err|passed|failed|Error|illegal" | cut -c1-250 | tail
timeout 120 python tools/selfplay_generation.py tictactoe 256 /tmp/g0 > gpurun_out/gen_ttt.json 2> gpurun_out/gen_ttt.err; echo rc=$?; tail -c 700 gpurun_out/gen_ttt.json; tail -3 gpurun_out/gen_ttt.err
timeout 600 python tools/selfplay_generation.py connect4 4096 /tmp/g1 > gpurun_out/gen_c4.json 2> gpurun_out/gen_c4.err; echo rc=$?; tail -c 900 gpurun_out/gen_c4.json; tail -3 gpurun_out/gen_c4.err
