for v in base pw_noload; do ABOPT_LIB=ab_opt_b200/_lib/variants/$v/libabopt_b200.so python scripts/kbench.py --config c2 --steps 6 --tag $v 2>&1 | tail -1; done > gpurun_out/r02_kbench_e.log
python scripts/kbench.py --config c2 --steps 6 --tag new 2>&1 | tail -1 >> gpurun_out/r02_kbench_e.log
ABOPT_LIB=ab_opt_b200/_lib/variants/base/libabopt_b200.so python scripts/kbench.py --config c2 --steps 6 --tag base 2>&1 | tail -1 >> gpurun_out/r02_kbench_e.log
python scripts/kbench.py --config c2 --steps 6 --tag new 2>&1 | tail -1 >> gpurun_out/r02_kbench_e.log
cat gpurun_out/r02_kbench_e.log
