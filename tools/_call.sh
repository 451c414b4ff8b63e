timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_index.py tests/test_gpu_scale.py -m gpu -q -x -p no:cacheprovider -k "not 5k and not configs3" > gpurun_out/r02k_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02k_tests.log
tools/gpu_round.sh r02k bench
tools/gpu_profile.sh r02k a
