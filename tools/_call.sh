timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02i_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02i_tests.log
tools/gpu_round.sh r02i bench
python bench.py --steps 20 --warmup 3 --no-ksw --no-regime-b --no-cpu-baseline > gpurun_out/r02i_bench_20steps.json 2>/dev/null; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench_20steps.json'))
print("20 steps:", d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['two_context_pool'])
PY
tools/gpu_profile.sh r02i a
