#!/bin/bash
# ncu evidence for profiles/: launch list of the bench's timed region, full captures of the regime-A kernels and of the ksw kernel.
TAG=${1:-x}
mkdir -p gpurun_out
export ND_NCU_RANGE=1
B="python bench.py --steps 3 --warmup 3 --no-ksw --no-regime-b --no-cpu-baseline"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv $B > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -c 60 -f -o gpurun_out/${TAG}_regimeA $B > gpurun_out/${TAG}_regimeA.log 2>&1
echo "regime A full exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ksw -c 4 -f -o gpurun_out/${TAG}_ksw python tools/ksw_only.py 20000 > gpurun_out/${TAG}_ksw.log 2>&1
echo "ksw full exit $?"
ls -la gpurun_out
