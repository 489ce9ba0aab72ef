set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02n_gputests.txt 2>&1; tail -3 gpurun_out/r02n_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -2 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --serial > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_core|k_legacy|k_warp|k_lanczos|k_rotate|k_flip|k_distort" -c 11 -f -o gpurun_out/r02_step python tools/prof_step.py 4096 1 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
LFX_CORE_TIMING=1 python tools/prof_step.py 4096 1 2> gpurun_out/r02_k_core_phase_split.txt >/dev/null
python tools/bench_ops.py --batch 4096 --size 256 > gpurun_out/r02_ops_256.json 2>/dev/null
python tools/bench_ops.py --batch 256 --size 1024 > gpurun_out/r02_ops_1024.json 2>/dev/null
python tools/bench_jpeg.py > gpurun_out/r02_jpeg.json 2>/dev/null
ls -la gpurun_out/r02_* | head -30
