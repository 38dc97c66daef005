mkdir -p gpurun_out
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --workload c5 --gpus 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02n_c5_8k_1024spp_8gpu.json 2> gpurun_out/scale_r02n_c5.err
echo rc=$?; cut -c1-260 gpurun_out/scale_r02n_c5_8k_1024spp_8gpu.json; tail -2 gpurun_out/scale_r02n_c5.err | cut -c1-200
