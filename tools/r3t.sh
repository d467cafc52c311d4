set -x
O=gpurun_out
for ns in 0 2000 3500 5000 6500 8000; do
  echo "== stagger $ns" >> $O/r3t.txt
  B200_CANNY_STAGGER_NS=$ns timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-bands --no-extras > $O/r3t_$ns.json 2>> $O/r3t.err
  python - <<PY >> $O/r3t.txt
import json
d=json.loads(open("gpurun_out/r3t_$ns.json").read()); print(d["value"], d["ms_per_step"], "alone launch_ms", d["roofline"]["launch_ms"])
PY
done
cat $O/r3t.txt
