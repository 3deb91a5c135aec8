N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo rc=$?; tail -c 600 gpurun_out/scale_$N.json
