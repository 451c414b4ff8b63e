export ND_NCU_RANGE=1
timeout 110 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-ksw --no-regime-b --no-cpu-baseline > gpurun_out/r02n_launches_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r02n_launches_bench.csv
