timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02n_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02n_tests.log
timeout 300 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench exit $?"; cat gpurun_out/r02n_bench_n1.json | head -c 1200
