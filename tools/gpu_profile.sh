#!/bin/bash
# ncu evidence for profiles/: launch list of the bench's timed region, full captures of the regime-A kernels and of the ksw kernel.
# Everything that travels back must stay well under 64 MiB: CSV exports are made on the box and large reports are dropped.
TAG=${1:-x}; WHAT=${2:-all}
mkdir -p gpurun_out
export ND_NCU_RANGE=1
B="python bench.py --steps 2 --warmup 3 --no-ksw --no-regime-b --no-cpu-baseline"
keep_small() { for f in "$@"; do if [ -f "$f" ] && [ $(stat -c %s "$f") -gt 25000000 ]; then rm -f "$f"; fi; done; }
if [ "$WHAT" = all ] || [ "$WHAT" = a ]; then
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv $B > gpurun_out/${TAG}_launches_bench.log 2>&1
  echo "launch list exit $?"
  timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -c 48 -f -o /tmp/${TAG}_regimeA $B > gpurun_out/${TAG}_regimeA.log 2>&1
  echo "regime A full exit $?"
  ncu -i /tmp/${TAG}_regimeA.ncu-rep --page raw --csv > gpurun_out/${TAG}_regimeA_raw.csv 2>/dev/null
  for k in k_sketch_fast4 k_chain_fill_warp k_probe_count k_rs_init; do
    ncu -i /tmp/${TAG}_regimeA.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
  done
fi
if [ "$WHAT" = all ] || [ "$WHAT" = k ]; then
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ksw -c 2 -f -o /tmp/${TAG}_ksw python tools/ksw_only.py 20000 > gpurun_out/${TAG}_ksw.log 2>&1
  echo "ksw full exit $?"
  ncu -i /tmp/${TAG}_ksw.ncu-rep --page raw --csv > gpurun_out/${TAG}_ksw_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_ksw.ncu-rep --page source --csv -c 1 > gpurun_out/${TAG}_src_ksw.csv 2>/dev/null
  cp /tmp/${TAG}_ksw.ncu-rep gpurun_out/ 2>/dev/null
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ksw -c 6 -f -o /tmp/${TAG}_kswb python tools/ksw_regimeb.py > gpurun_out/${TAG}_kswb.log 2>&1
  echo "ksw regime-B full exit $?"
  ncu -i /tmp/${TAG}_kswb.ncu-rep --page raw --csv > gpurun_out/${TAG}_kswb_raw.csv 2>/dev/null
fi
keep_small gpurun_out/*.ncu-rep gpurun_out/*.csv
du -sh gpurun_out; ls -la gpurun_out | tail -30
