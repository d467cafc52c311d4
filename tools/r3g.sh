set -x
O=gpurun_out
for k in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --kind $k > $O/r3g_bench_k$k.json 2>> $O/r3g.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o $O/r3g_front3_9f python bench.py --frames 9 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r3g_ncu.log 2>&1
python - <<'PY'
import json
for k in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r3g_bench_k{k}.json").read())
        print("kind",k, d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"], d["roofline"]["stages"])
        print(d.get("content")); print(d.get("bgr")); print(d.get("latency"))
    except Exception as e: print(k,"ERR",e)
PY
tail -5 $O/r3g.err
