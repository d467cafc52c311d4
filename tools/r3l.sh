set -x
O=gpurun_out
for v in R96 R88 R80 R72; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in -1 1 0; do timeout 300 python tools/stage_times.py --frames 63 --kind $k > $O/r3l_${v}_k$k.json 2>> $O/r3l.err; done
done
cp ab/libR96.so canny_edge_b200/libcanny_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3l_*_k*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], "pipeline", d["pipeline_ms"], round(d["pipeline_Mpix_s"]/1e3,1), "front", d["stages"]["front"]["ms"], "link", d["stages"]["ccl_local"]["ms"], "resolve", d["stages"]["ccl_final"]["ms"])
PY
