set -x
O=gpurun_out
rm -f $O/r3l_*
for r in 1 2; do
for v in ${AB_LIBS:-R96 R104 R112}; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in -1 1 0; do timeout 300 python tools/stage_times.py --frames 63 --kind $k > $O/r3l_${v}_k${k}_r$r.json 2>> $O/r3l.err; done
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras > $O/r3l_${v}_bench_r$r.json 2>> $O/r3l.err
done
done
cp ab/libR96.so canny_edge_b200/libcanny_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3l_*_k*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], "pipeline", d["pipeline_ms"], round(d["pipeline_Mpix_s"]/1e3,1), "front", d["stages"]["front"]["ms"], "link", d["stages"]["ccl_local"]["ms"], "resolve", d["stages"]["ccl_final"]["ms"])
for f in sorted(glob.glob("gpurun_out/r3l_*_bench_r*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], d["value"], d["ms_per_step"], d["roofline"]["launch_ms"])
PY
