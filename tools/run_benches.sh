#!/bin/bash
# all bench workloads on one GPU -> gpurun_out/<tag>_bench_*.json + a short summary on stdout
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_seg.json 2> gpurun_out/${tag}_bench_seg.err; tail -3 gpurun_out/${tag}_bench_seg.err
python bench.py --workload cls > gpurun_out/${tag}_bench_cls.json 2> gpurun_out/${tag}_bench_cls.err; tail -3 gpurun_out/${tag}_bench_cls.err
python bench.py --workload knn_ds > gpurun_out/${tag}_bench_knn_ds.json 2> gpurun_out/${tag}_bench_knn_ds.err; tail -3 gpurun_out/${tag}_bench_knn_ds.err
if [ "$2" == "ref" ]; then python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2>&1; fi
python - <<PY
import json
for w in ("seg","cls"):
    try:
        d=json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "launches/step", d["gpu_launches"]//d["steps"], d["clocks"])
        print("  north:", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k not in ("note","peak_source")})
        print("  linear:", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline_linear"].items() if k not in ("note",)})
        print("  kernels:", d["kernels_ms_per_step"])
        print("  cpu:", d.get("cpu_baseline"))
    except Exception as e: print(w, "ERR", e)
try:
    d=json.load(open("gpurun_out/${tag}_bench_knn_ds.json")); print(json.dumps(d["results"])); print(d.get("cpu_baseline"))
except Exception as e: print("knn_ds ERR", e)
try: print(open("gpurun_out/${tag}_bench_ref.json").read()[-700:])
except Exception: pass
PY
