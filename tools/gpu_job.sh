# scratch driver of one gpurun call (rewritten per call; see tools/README.md)
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --games 32768 --pool-fraction 0.45 --no-cpu-baseline --no-extras > gpurun_out/r02_scale_gomoku_4gpu_131072.json 2> gpurun_out/scale4.err; echo rc=$?
tail -5 gpurun_out/scale4.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_scale_gomoku_4gpu_131072.json') if l.startswith('{')][-1])
print('N=4', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config'].get('games_per_gpu'), d['clocks'], d['hbm_bytes'], d['slot_pool'])
print(json.dumps(d.get('e2e_generation'))[:900])
"
