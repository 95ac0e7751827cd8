mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
bash tools/sanitize.sh gpurun_out/r2b 1gpu
for cfg in c2 c3 c4 c5; do timeout 400 python bench.py --config $cfg > gpurun_out/r2b/bench_${cfg}.json 2> gpurun_out/r2b/bench_${cfg}.err; tail -c 400 gpurun_out/r2b/bench_${cfg}.err; done
python - <<'PY'
import json
for c in ("c2","c3","c4","c5"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "ms/plan", round(d["ms_per_step"],4), "p50", round(d["ms_per_step_p50"],4), "value", f'{d["value"]:.4g}', "e2e", f'{d["e2e"]["value"]:.4g}', d["gpu_config"]["k2_mapping"], {k: round(v,1) for k,v in d.get("kernels_us",{}).items()})
    except Exception as e: print(c, "ERR", e)
PY
