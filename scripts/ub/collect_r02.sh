# round-2 evidence: ncu full capture (-> DRAM traffic of the pair kernel), bench lines, launch list, GPU test log -> gpurun_out/r02_final_*
set -x
ncu --set full --clock-control none --import-source on --kernel-name regex:'pair_stream|attn_logits|aggr_persist|outT_tail|proj_ts|ctx_delta' --launch-skip 64 --launch-count 32 -o gpurun_out/r02_final_prof -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_final_ncu.log 2>&1
ncu -i gpurun_out/r02_final_prof.ncu-rep --page raw --csv > gpurun_out/r02_final_raw.csv 2>/dev/null
python scripts/pair_traffic.py gpurun_out/r02_final_raw.csv gpurun_out/pair_kernel_traffic.json "ncu --set full --clock-control none, profiles/r02_final_ncu_full_summary.json" && cp gpurun_out/pair_kernel_traffic.json profiles/pair_kernel_traffic.json
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_final_pytest_gpu.log
python bench.py > gpurun_out/r02_final_bench_c2.json 2> gpurun_out/r02_final_bench_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
for c in c3 c4 c5 f1; do python bench.py --config $c --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_final_bench_$c.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches_c2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-eager > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'pair_embed_tc' --launch-skip 1 --launch-count 1 -o gpurun_out/r02_final_prof_f1 -f python bench.py --config f1 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_final_ncu_f1.log 2>&1
tail -3 gpurun_out/r02_final_pytest_gpu.log; head -c 400 gpurun_out/r02_final_bench_c2.json
