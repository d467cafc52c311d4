set -x
O=gpurun_out
for r in 1 2 3; do
for v in A B; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in 0 1; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > $O/r3j_${v}_k${k}_r$r.json 2>> $O/r3j.err
  done
done
done
cp ab/libB.so canny_edge_b200/libcanny_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q 2>&1 | tail -4 > $O/r3j_tests.log
for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r3j_pdl$v.json 2>> $O/r3j.err; done
cat $O/r3j_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3j_?_k*_r*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split('/')[-1], d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"])
    except Exception as e: print(f,"ERR",e)
PY
cat $O/r3j_pdl*.json | cut -c1-300
