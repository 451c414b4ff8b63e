timeout 2700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02l_tests.log 2>&1; echo "tests exit $?"; tail -30 gpurun_out/r02l_tests.log
