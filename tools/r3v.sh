set -x
O=gpurun_out
rm -f $O/r3v_*
cp ab/libB.so canny_edge_b200/libcanny_b200.so
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r3v_tests.log; cat $O/r3v_tests.log
for r in 1 2; do
for v in A B; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in 0 1; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > $O/r3v_${v}_k${k}_r$r.json 2>> $O/r3v.err
  done
  timeout 300 python tools/stage_times.py --frames 63 --kind -1 > $O/r3v_${v}_photo_r$r.json 2>> $O/r3v.err
done
done
cp ab/libB.so canny_edge_b200/libcanny_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3v_?_k*_r*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"])
for f in sorted(glob.glob("gpurun_out/r3v_?_photo_r*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], "pipeline", d["pipeline_ms"], round(d["pipeline_Mpix_s"]/1e3,1), "front", d["stages"]["front"]["ms"])
PY
