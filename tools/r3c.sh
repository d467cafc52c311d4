set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_bgr_cli.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > $O/r3c_tests.log
cat $O/r3c_tests.log
timeout 300 python tools/bgr_probe.py > $O/r3c_bgr_fused.json 2>> $O/r3c.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py > $O/r3c_bgr_separate.json 2>> $O/r3c.err
timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --sigma 1.4 --steps 20 > $O/r3c_bgr8k_fused.json 2>> $O/r3c.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --sigma 1.4 --steps 20 > $O/r3c_bgr8k_separate.json 2>> $O/r3c.err
for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r3c_pdl$v.json 2>> $O/r3c.err; done
timeout 300 python tools/pdl_probe.py > $O/r3c_pdl_default.json 2>> $O/r3c.err
cat $O/r3c_bgr*.json $O/r3c_pdl*.json; tail -5 $O/r3c.err
