# scratch driver for one gpurun call: Connect4 launch list at HEAD
timeout 300 python bench.py --config connect4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c4_plain.log 2>&1; echo plain rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 80 --csv --log-file gpurun_out/launches_c4_head.csv python bench.py --config connect4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c4_ncu.log 2>&1; echo ncu rc=$?
