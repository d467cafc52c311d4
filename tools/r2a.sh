set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_tests.log
B200_CANNY_LOCAL_LINK=1 timeout 300 python tests/scripts/local_link_check.py > gpurun_out/r2a_ll.log 2>&1; echo "ll exit $?" >> gpurun_out/r2a_ll.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
B200_CANNY_LOCAL_LINK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_ll.json 2> gpurun_out/r2a_bench_ll.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front2 -c 1 -f -o gpurun_out/r2a_front2 python bench.py --frames 4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2a_ncu.log 2>&1
ls -la gpurun_out
