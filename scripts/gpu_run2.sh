set -x
ARIA_TEST_POISON=1 timeout 1200 python -m pytest tests/test_gpu_lora.py tests/test_gpu_parity_full.py tests/test_gpu_dropin.py -q -s > gpurun_out/r02_newtests.log 2>&1
tail -60 gpurun_out/r02_newtests.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_dropin.py > gpurun_out/r02_pytest_all.log 2>&1
tail -8 gpurun_out/r02_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -5 gpurun_out/r02_smoke.log
