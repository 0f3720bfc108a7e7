# A/B of whole-bench runs: variants given as arguments ("new" = the default library)
out=gpurun_out/${OUT:-r02_bench_ab.log}; : > $out
for i in 1 2; do
for v in "$@"; do
  if [ $v = new ]; then unset ABOPT_LIB; else export ABOPT_LIB=ab_opt_b200/_lib/variants/$v/libabopt_b200.so; fi
  python bench.py --no-cpu-baseline --no-gpu-eager 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],1), d.get('ms_each_step'), 'e2e', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], {k:v['ms'] for k,v in d['kernel_breakdown_ms_per_sample'].items()})
" >> $out
done; done
cat $out
