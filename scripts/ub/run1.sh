for v in pw_nocomp pw_noalpha pw_nocomp_noalpha; do ABOPT_LIB=ab_opt_b200/_lib/variants/$v/libabopt_b200.so python scripts/kbench.py --config c2 --tag $v 2>&1 | tail -1; done > gpurun_out/r02_kbench_variants.log
ncu --set full --clock-control none --import-source on --kernel-name regex:pair_stream -c 3 -o gpurun_out/r02_prof_pair -f python scripts/profile_step.py --config c2 --steps 1 > gpurun_out/r02_ncu_pair.log 2>&1
cat gpurun_out/r02_kbench_variants.log; tail -3 gpurun_out/r02_ncu_pair.log
