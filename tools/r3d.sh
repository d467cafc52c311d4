set -x
O=gpurun_out
# A = old phase 3b, B = two entries per lane, branch-free
for r in 1 2; do
for v in A B; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  for k in 0 1; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind $k > $O/r3d_${v}_k${k}_r$r.json 2>> $O/r3d.err
  done
done
done
cp ab/libB.so canny_edge_b200/libcanny_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bgr_cli.py -m gpu -x -q 2>&1 | tail -5 > $O/r3d_tests.log
timeout 300 python tools/bgr_probe.py > $O/r3d_bgr_fused.json 2>> $O/r3d.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py > $O/r3d_bgr_separate.json 2>> $O/r3d.err
timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --sigma 1.4 --steps 20 > $O/r3d_bgr8k_fused.json 2>> $O/r3d.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --sigma 1.4 --steps 20 > $O/r3d_bgr8k_separate.json 2>> $O/r3d.err
for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r3d_pdl$v.json 2>> $O/r3d.err; done
timeout 600 python tools/chunk_sweep.py > $O/r3d_chunk_sweep.txt 2>> $O/r3d.err
./tools/probes/blur_probe > $O/r3d_blur_probe.txt 2>> $O/r3d.err
cat $O/r3d_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3d_?_k*_r*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split('/')[-1], d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"])
    except Exception as e: print(f,"ERR",e)
PY
cat $O/r3d_bgr*.json $O/r3d_pdl*.json $O/r3d_chunk_sweep.txt $O/r3d_blur_probe.txt; tail -5 $O/r3d.err
