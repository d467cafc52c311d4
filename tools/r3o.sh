set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r3o_tests.log; cat $O/r3o_tests.log
for r in 1 2; do for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r3o_pdl${v}_r$r.json 2>> $O/r3o.err; done; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-bands > $O/r3o_bench.json 2>> $O/r3o.err
cat $O/r3o_pdl*.json | cut -c1-330
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3o_bench.json").read()); print(d["value"], d["ms_per_step"], d["latency"], d["e2e"]["value"], d["e2e"].get("pcie"))
PY
tail -3 $O/r3o.err
