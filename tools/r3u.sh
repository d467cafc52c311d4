set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_bgr_cli.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 > $O/r3u_tests.log; cat $O/r3u_tests.log
timeout 600 python tests/scripts/fuzz_bgr.py 30 40 2>&1 | tail -1
for r in 1 2; do timeout 300 python tools/bgr_probe.py > $O/r3u_bgr_r$r.json 2>> $O/r3u.err; done
timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --steps 20 > $O/r3u_bgr8k.json 2>> $O/r3u.err
cat $O/r3u_bgr*.json | cut -c1-300
