# final verification at HEAD: GPU suite at the default settings, then the C2 / C3 bench lines (no CPU / eager legs)
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_head_pytest_gpu.log
cat gpurun_out/r02_head_pytest_gpu.log
timeout 100 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_head_bench_c2.json 2> gpurun_out/r02_head_bench_c2.err
head -c 300 gpurun_out/r02_head_bench_c2.json
timeout 100 python bench.py --config c3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_head_bench_c3.json 2> gpurun_out/r02_head_bench_c3.err
head -c 300 gpurun_out/r02_head_bench_c3.json
