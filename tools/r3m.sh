set -x
O=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front3_kernel -s 5 -c 1 -f -o $O/r02_front3_bgr_9f python tools/bgr_probe.py --frames 9 --steps 1 > $O/r02_ncu_front_bgr.log 2>&1
tail -2 $O/r02_ncu_front_bgr.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
