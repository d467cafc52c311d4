#!/bin/bash
# N independent processes (one per GPU) running the host-path probe at the same time: how does the shared host (DRAM, PCIe root,
# cores) limit the aggregate?   usage: tools/e2e_multi.sh N [env assignments...]
N=$1; shift
for g in $(seq 0 $((N-1))); do
  env CUDA_VISIBLE_DEVICES=$g FRAMES=128 CHUNKS=8 "$@" python tools/e2e_probe.py > /tmp/e2e_$g.log 2>&1 &
done
wait
python - <<PY
import json,glob
tot=0
for f in sorted(glob.glob('/tmp/e2e_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); v=list(d['Gpix_s_by_chunk_frames'].values())[0]; tot+=v
    except Exception as e:
        print(f, 'failed', open(f).read()[-300:])
print("N=$N", "$*", "aggregate Gpix/s", round(tot,1), "per GPU", round(tot/$N,2))
PY
rm -f /tmp/e2e_*.log
