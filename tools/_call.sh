timeout 1200 python -m pytest tests/test_gpu_ovlsort.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02g_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r02g_tests.log
timeout 900 python tools/ovlsort_bench.py 30000 > gpurun_out/r02g_ovlsort_bench.json 2> gpurun_out/r02g_ovlsort_bench.err; cat gpurun_out/r02g_ovlsort_bench.json; tail -3 gpurun_out/r02g_ovlsort_bench.err
