mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02n.log 2>&1; tail -3 gpurun_out/pytest_r02n.log
python bench.py > gpurun_out/bench_r02n_c4_1gpu.json 2> gpurun_out/bench_r02n.err || tail -5 gpurun_out/bench_r02n.err
cut -c1-300 gpurun_out/bench_r02n_c4_1gpu.json
python tools/profile_run.py --workload c4 --width 3840 --height 2160 --spp 8 > gpurun_out/profile_run_r02n.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -f -o gpurun_out/r02n_k_render python tools/profile_run.py --workload c4 --width 3840 --height 2160 --spp 8 > gpurun_out/ncu_full_r02n.log 2>&1
ls -la gpurun_out/*.ncu-rep
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_r02n_launches_bench_py.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r02n.log 2>&1
tail -2 gpurun_out/ncu_launches_r02n.log | cut -c1-200
