timeout 2000 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02e_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02e_tests.log
python tools/ksw_only.py 20000 > gpurun_out/r02e_ksw_only.log 2>&1; tail -1 gpurun_out/r02e_ksw_only.log | cut -c1-300
NDG_ALN_DEBUG=1 python tools/ksw_regimeb.py > gpurun_out/r02e_regimeb.log 2>&1; grep "launch 0" gpurun_out/r02e_regimeb.log | cut -c1-300
python bench.py --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print("bench:", d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.05})
print(d['ksw2_extd2']); print(d['regime_b']); print(d['step2'])
PY
tools/gpu_profile.sh r02e k
