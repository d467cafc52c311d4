set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2d_tests.log
for k in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > gpurun_out/r2d_bench_f3_k$k.json 2>> gpurun_out/r2d.err
B200_CANNY_FRONT=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > gpurun_out/r2d_bench_f2_k$k.json 2>> gpurun_out/r2d.err
done
cat gpurun_out/r2d_tests.log
python - <<'PY'
import json
for k in (0,1):
    for f in ("f3","f2"):
        try:
            d=json.loads(open(f"gpurun_out/r2d_bench_{f}_k{k}.json").read())
            print(f, "kind",k, d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["stages"]["front"]["ms"])
        except Exception as e: print(f,k,"ERR",e)
PY
