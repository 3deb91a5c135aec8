# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:headconv_wide -s 10 -c 1 -o gpurun_out/prof_headwide2 -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_hw.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 10 -c 1 -o gpurun_out/prof_stemmma2 -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_st.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:res_block -s 52 -c 1 -o gpurun_out/prof_block_c4 -f python tools/quick_net_bench.py connect4 4096 > gpurun_out/ncu_bk.log 2>&1; echo ncu rc=$?
timeout 900 python bench.py --config connect4 --moves 8 --no-cpu-baseline --no-extras > gpurun_out/r02_moves_connect4.json 2> gpurun_out/m1.err; echo rc=$?
timeout 900 python bench.py --config connect4 --moves 8 --no-cpu-baseline --no-extras --eval-cache 8388608 > gpurun_out/r02_moves_connect4_cache.json 2> gpurun_out/m2.err; echo rc=$?
tail -n 3 gpurun_out/m1.err gpurun_out/m2.err
python -c "
import json
for f in ('r02_moves_connect4','r02_moves_connect4_cache'):
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, d['value'], d['ms_per_step'], d['sims_per_eval'], d['nn_evals_per_s'], d['positions_per_s'], d.get('eval_cache'))
"
