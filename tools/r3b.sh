# A/B of front3 variants: AB_LIBS="A B C D"; parity suite on the default build first
set -x
O=gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > $O/r3b_tests.log
cat $O/r3b_tests.log 2>/dev/null
for r in 1 2; do
for v in ${AB_LIBS:-A B C D}; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in 0 1; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > $O/r3b_${v}_k${k}_r$r.json 2>> $O/r3b.err
  done
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3b_*_k*_r*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split('/')[-1], d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"])
    except Exception as e: print(f,"ERR",e)
PY
