# scratch driver for one gpurun call: ncu --set full capture of the C128->C32 head convolution (conv_board_kernel<32, pair>)
timeout 600 ncu --set full --clock-control none --import-source on -k conv_board_kernel -s 109 -c 1 -o gpurun_out/prof_conv32_r1 -f python bench.py --steps 3 --warmup 3 --presearch 8 --no-cpu-baseline > gpurun_out/conv32_ncu.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/prof_conv32_r1.ncu-rep; tail -2 gpurun_out/conv32_ncu.log | cut -c1-200
