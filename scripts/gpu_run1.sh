set -x
nvidia-smi -L
python scripts/stress_small_groups.py 30 > gpurun_out/r02_stress1.log 2>&1
tail -60 gpurun_out/r02_stress1.log
ARIA_TEST_POISON=1 timeout 300 python -m pytest tests/test_zz_gpu_round1_unverified.py tests/test_gpu_lora.py -q --runxfail -x -k "not two_devices" > gpurun_out/r02_zz_poison.log 2>&1
tail -40 gpurun_out/r02_zz_poison.log
ARIA_TEST_POISON=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_poison.log 2>&1
tail -15 gpurun_out/r02_pytest_poison.log
