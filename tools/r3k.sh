set -x
O=gpurun_out
for k in -1 1 0; do timeout 300 python tools/stage_times.py --frames 63 --kind $k > $O/r3k_stage_k$k.json 2>> $O/r3k.err; done
cat $O/r3k_stage_k*.json | cut -c1-900
