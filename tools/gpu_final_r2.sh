#!/bin/bash
# round 2 evidence run on one B200: stamp trace of the dataflow kernel, every one-clip schedule side by side, bench lines of configs[1],
# configs[2] (60 s clip) and 32 clips per GPU, the reference arm
mkdir -p gpurun_out
FLOW_TRACE_TAG=final FLOW_TRACE_CTA=17 FLOW_TRACE_T0=150 FLOW_TRACE_T1=170 timeout 200 python tools/flow_trace.py 1 v > gpurun_out/flow_trace_final.txt 2>&1
timeout 300 python tools/flow_check.py > gpurun_out/flow_check_final.txt 2>&1; tail -12 gpurun_out/flow_check_final.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_final_c1.json 2> gpurun_out/bench_final_c1.err
timeout 600 python bench.py --frames 1500 --steps 5 --warmup 3 --no-cpu-baseline --large-clips 0 > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err
timeout 600 python bench.py --batch 32 --frames 200 --steps 5 --warmup 3 --no-cpu-baseline --large-clips 0 > gpurun_out/bench_final_b32.json 2> gpurun_out/bench_final_b32.err
python - <<'PY'
import json
for n in ("c1", "c2", "b32", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/bench_final_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 1), d.get("us_per_ode_step"), d.get("roofline", {}).get("frac"), d.get("e2e", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(n, "failed", e)
PY
