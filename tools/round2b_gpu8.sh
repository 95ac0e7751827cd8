# 8 GPUs: BASELINE C5 at its 8 GPUs and C2 at 8 GPUs (strong scaling, straggler check of the rank-0 clock thread)
mkdir -p gpurun_out/r2b
run() { cfg=$1; n=$2; shift 2; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --config $cfg "$@" > gpurun_out/r2b/bench_${cfg}_n${n}.json 2> gpurun_out/r2b/bench_${cfg}_n${n}.err; tail -c 200 gpurun_out/r2b/bench_${cfg}_n${n}.err; }
run c5 8
run c2 8
python - <<'PY'
import json
for f in ("c5_n8","c2_n8"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/plan", round(d["ms_per_step"],4), "p50", round(d["ms_per_step_p50"],4), "value", f'{d["value"]:.4g}', "e2e", f'{d["e2e"]["value"]:.4g}', d["gpu_config"]["k2_mapping"], d.get("correctness"), d["gpu_config"]["ms_per_step_p10_p50_p90_max"])
    except Exception as e: print(f, "ERR", e)
PY
