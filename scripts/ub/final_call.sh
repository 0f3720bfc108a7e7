# final verification at HEAD: GPU suite at the default settings, then bench lines without the CPU / eager legs
set -x
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_head_pytest_gpu.log
cat gpurun_out/r02_head_pytest_gpu.log
for c in "$@"; do
  timeout 40 python bench.py --config $c --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_head_bench_$c.json 2> gpurun_out/r02_head_bench_$c.err
  head -c 300 gpurun_out/r02_head_bench_$c.json
done
