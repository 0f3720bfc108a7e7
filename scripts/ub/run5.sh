python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k optimize_replayed 2>&1 | tail -5
ncu --set full --clock-control none --import-source on --kernel-name regex:attn_logits -c 2 -o gpurun_out/r02_prof_logits_pp -f python scripts/profile_step.py --config c2 --steps 1 > gpurun_out/r02_ncu_logits.log 2>&1
tail -2 gpurun_out/r02_ncu_logits.log
