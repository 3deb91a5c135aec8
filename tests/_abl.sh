timeout 600 python -m pytest tests/test_net_gpu.py -x -q 2>&1 | tail -3
for d in 0 0; do timeout 120 python tests/quick_net_bench.py gomoku 16384 2>&1 | grep -E "batch 16384|per conv" | tail -7; done
