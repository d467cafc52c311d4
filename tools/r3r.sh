set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q 2>&1 | tail -4 > $O/r3r_tests.log; cat $O/r3r_tests.log
for r in 1 2; do
for v in ${AB_LIBS:-A B}; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in -1 1 0; do timeout 300 python tools/stage_times.py --frames 63 --kind $k > $O/r3r_${v}_k${k}_r$r.json 2>> $O/r3r.err; done
  timeout 300 python tools/pdl_probe.py > $O/r3r_${v}_lat_r$r.json 2>> $O/r3r.err
done
done
cp ab/libB.so canny_edge_b200/libcanny_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3r_*_k*_r*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], "pipeline", d["pipeline_ms"], round(d["pipeline_Mpix_s"]/1e3,1), "front", d["stages"]["front"]["ms"], "link", d["stages"]["ccl_local"]["ms"], "resolve", d["stages"]["ccl_final"]["ms"])
for f in sorted(glob.glob("gpurun_out/r3r_*_lat_r*.json")):
    d=json.loads(open(f).read()); print(f.split('/')[-1], {k:v["us_per_frame"] for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/r3r.err
