L=grok_alpha_zero_b200/libgaz_b200.so
cp $L /tmp/head.so
timeout 300 python -m pytest tests/test_net_gpu.py -q -s -k "test_net_matches_fp32_oracle" 2>&1 | grep -E "err|passed|failed"
for rep in 1 2 3; do
  for v in head prev; do
    if [ $v = head ]; then cp /tmp/head.so $L; else cp tests/_emul/libgaz_prev.so $L; fi
    echo "== $v rep $rep"; timeout 200 python tests/quick_net_bench.py gomoku 16384 2>&1 | grep -E "batch 16384|per conv" | tail -2
  done
done
cp /tmp/head.so $L
