set -x
mkdir -p gpurun_out
bash tools/scale.sh 2
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_2.json') if l.startswith('{')][-1])
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['clocks'])
print(json.dumps(d.get('configs', d.get('e2e_generation')))[:1500])
"
tail -5 gpurun_out/scale_2.err
