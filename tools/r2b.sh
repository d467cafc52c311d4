set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "rc $?" >> gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err
tail -5 gpurun_out/r2b_bench.err
