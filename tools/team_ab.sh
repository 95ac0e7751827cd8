#!/bin/bash
# A/B of the two K2 mappings on the contact configs: GPU tests under both, then the bench lines (kernel times inside)
mkdir -p gpurun_out/team
MPPIB_K2_TEAM=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/team/tests_team.log
MPPIB_K2_TEAM=0 timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5 > gpurun_out/team/tests_thread.log
for cfg in c3 c4 c5; do
  for team in 0 1; do
    MPPIB_K2_TEAM=$team timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/team/bench_${cfg}_team${team}.json 2> gpurun_out/team/bench_${cfg}_team${team}.err
  done
done
tail -15 gpurun_out/team/tests_team.log; tail -3 gpurun_out/team/tests_thread.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/team/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("kernels_us") or d.get("kernel_us") or {k: v for k, v in d.items() if "kernel" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
