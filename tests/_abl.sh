timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q 2>&1 | tail -5
L=grok_alpha_zero_b200/libgaz_b200.so
cp $L /tmp/head.so
for rep in 1 2; do
  for v in head prev; do
    if [ $v = head ]; then cp /tmp/head.so $L; else cp tests/_emul/libgaz_prev.so $L; fi
    echo "== $v rep $rep"; timeout 200 python tests/quick_net_bench.py gomoku 16384 2>&1 | grep -E "batch 16384|per conv" | tail -2
  done
done
cp /tmp/head.so $L
