timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_e.log
cat gpurun_out/r02_pytest_e.log
for v in base new; do
  if [ $v = new ]; then unset ABOPT_LIB; else export ABOPT_LIB=ab_opt_b200/_lib/variants/$v/libabopt_b200.so; fi
  timeout 300 python scripts/kbench.py --config c2 --steps 6 --tag $v 2>&1 | tail -1
  timeout 300 python scripts/kbench.py --config c4 --steps 6 --tag $v-c4 2>&1 | tail -1
done > gpurun_out/r02_kbench_f.log
cat gpurun_out/r02_kbench_f.log
