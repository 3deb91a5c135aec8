# scratch driver for one gpurun call: ncu --set full capture of the tree kernels mid-search (Gomoku, 16384 games)
timeout 300 python bench.py --steps 3 --warmup 3 --presearch 232 --no-cpu-baseline > gpurun_out/tree_plain.log 2>&1; echo plain rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_select|k_expand' -s 476 -c 4 -o gpurun_out/prof_tree_r1 -f python bench.py --steps 3 --warmup 3 --presearch 232 --no-cpu-baseline > gpurun_out/tree_ncu.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/tree_ncu.log
ls -la gpurun_out/prof_tree_r1.ncu-rep
