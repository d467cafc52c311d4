set -x
O=gpurun_out
timeout 1200 python tests/scripts/fuzz_more.py 100 250 > $O/r3q_fuzz.log 2>&1; tail -2 $O/r3q_fuzz.log
timeout 900 python tests/scripts/fuzz_bgr.py 0 25 > $O/r3q_fuzz_bgr.log 2>&1; tail -2 $O/r3q_fuzz_bgr.log
