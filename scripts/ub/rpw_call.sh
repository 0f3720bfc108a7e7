# one gpurun call: A/B of the rows-per-warp settings (one process), then the GPU suite with the candidate setting
set -x
mkdir -p gpurun_out
timeout 150 python scripts/ub/rpw_ab.py > gpurun_out/rpw_ab.jsonl 2> gpurun_out/rpw_ab.err
cat gpurun_out/rpw_ab.jsonl
ABOPT_RPW_MIXER=2 ABOPT_RPW_HEADS=4 timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/rpw_pytest.log
cat gpurun_out/rpw_pytest.log
