set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r3f_tests.log
cat $O/r3f_tests.log
for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r3f_pdl$v.json 2>> $O/r3f.err; done
for k in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --kind $k > $O/r3f_bench_k$k.json 2>> $O/r3f.err
done
cat $O/r3f_pdl*.json
python - <<'PY'
import json
for k in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r3f_bench_k{k}.json").read())
        print("kind",k, d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"], d["roofline"]["stages"])
        print(d.get("content")); print(d.get("bgr")); print(d.get("latency")); print(d.get("weak"))
    except Exception as e: print(k,"ERR",e)
PY
tail -5 $O/r3f.err
