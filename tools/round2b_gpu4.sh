# 4 GPUs: the GPU test suite (its 2-GPU tests run here), BASELINE C4 at its 4 GPUs (+ the NCCL exchange for reference), C2 at 4 GPUs
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
run() { cfg=$1; n=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $cfg "$@" > gpurun_out/r2b/bench_${cfg}_n${n}.json 2> gpurun_out/r2b/bench_${cfg}_n${n}.err; tail -c 300 gpurun_out/r2b/bench_${cfg}_n${n}.err; }
run c4 4
run c3 2
run c2 4
python - <<'PY'
import json
for f in ("c4_n4","c3_n2","c2_n4"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/plan", round(d["ms_per_step"],4), "p50", round(d["ms_per_step_p50"],4), "value", f'{d["value"]:.4g}', "e2e", f'{d["e2e"]["value"]:.4g}', d["gpu_config"]["k2_mapping"], d.get("correctness"))
    except Exception as e: print(f, "ERR", e)
PY
