export ND_NCU_RANGE=1
timeout 60 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_group_filter|k_sort_fast|k_chain_fill_warp' -c 6 -f -o /tmp/r02n_sort python bench.py --steps 1 --warmup 3 --no-ksw --no-regime-b --no-cpu-baseline > gpurun_out/r02n_sort_full.log 2>&1
echo "full exit $?"
ncu -i /tmp/r02n_sort.ncu-rep --page raw --csv > gpurun_out/r02n_sort_filter_ncu_full_raw.csv 2>/dev/null; wc -c gpurun_out/r02n_sort_filter_ncu_full_raw.csv
