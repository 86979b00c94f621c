mkdir -p gpurun_out
N=${1:-2}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py ) > gpurun_out/r2_dist_check_${N}gpu.log 2>&1; echo "dist_check rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/r2_dist_check_${N}gpu.log | head -20
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ); echo "bench n$N rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_bench_n{N}.json').read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
    s=d["sharded"]
    if "snip" in s: print("snip:", {k:v for k,v in s["snip"].items() if k not in ("what","bound_note")})
    for name,v in s["config5"].items():
        print(name, {k:x for k,x in v.items() if k not in ("levels","weights")})
        print("   ", {l:(round(x["one_gpu_us"],1), round(x["sharded_us"],1), round(x["speedup"],2), x["bit_identical"]) for l,x in v["levels"].items()})
    print("lost", d["lost"]["value"], d["lost"]["voc_mix"]["value"], d["lost"]["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
