# scratch driver for one gpurun call: pool relief on CUDA vs emulation, facade suite, Connect4 generation
timeout 600 python -m pytest tests/test_facade_gpu.py -q -x 2>&1 | tail -4
timeout 600 python tools/selfplay_generation.py connect4 4096 /tmp/g1 > gpurun_out/gen_c4b.json 2> gpurun_out/gen_c4b.err; echo rc=$?; tail -c 500 gpurun_out/gen_c4b.json; tail -3 gpurun_out/gen_c4b.err
