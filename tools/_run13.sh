mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "edge_cases or pixel_queue or golden_small or random_scenes_are_bit_exact and (1 or 2)" > gpurun_out/memcheck_r02r.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/memcheck_r02r.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02l_c4_4gpu.json 2> gpurun_out/scale_r02l_c4_4gpu.err
cut -c1-200 gpurun_out/scale_r02l_c4_4gpu.json
