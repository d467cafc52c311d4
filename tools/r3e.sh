set -x
O=gpurun_out
for r in 1 2; do
for v in A N; do
  cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
  B200_CANNY_TAIL_BANDS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras --kind 0 > $O/r3e_${v}_k0_r$r.json 2>> $O/r3e.err
done
done
cp ab/libA.so canny_edge_b200/libcanny_b200.so
for tb in 1 2 3 4; do for tc in 1 2 3; do
  echo "== tail_bands $tb tail_chunks $tc" >> $O/r3e_tail.txt
  B200_CANNY_TAIL_BANDS=$tb B200_CANNY_TAIL_CHUNKS=$tc timeout 300 python tools/chunk_sweep.py --frames 64,512 --chunks 0,6 2>> $O/r3e.err | grep frames >> $O/r3e_tail.txt
  [ $tb = 1 ] && break
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3e_?_k*_r*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split('/')[-1], d["value"], d["ms_per_step"], "front launch_ms", d["roofline"]["launch_ms"])
    except Exception as e: print(f,"ERR",e)
PY
cat $O/r3e_tail.txt; tail -5 $O/r3e.err
