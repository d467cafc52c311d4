set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r3h_tests.log
cat $O/r3h_tests.log
for k in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --kind $k > $O/r3h_bench_k$k.json 2>> $O/r3h.err
done
python - <<'PY'
import json
for k in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r3h_bench_k{k}.json").read())
        print("kind",k, d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"], d["roofline"]["stages"])
        print(d.get("content")); print(d.get("latency"))
    except Exception as e: print(k,"ERR",e)
PY
tail -5 $O/r3h.err
