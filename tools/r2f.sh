# quick A/B of the front kernel: parity subset + bench (shapes, noise) ; usage: tools/r2f.sh <tag>
tag=$1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_tests.log
for k in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > gpurun_out/${tag}_bench_k$k.json 2>> gpurun_out/${tag}.err
done
cat gpurun_out/${tag}_tests.log
python - <<PY
import json
for k in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/${tag}_bench_k{k}.json").read())
        print("kind",k, d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"], d["roofline"]["stages"])
    except Exception as e: print(k,"ERR",e)
PY
