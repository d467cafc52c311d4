set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tests/scripts/multigpu_bands_check.py --height 4096 --width 4096 --kind 1 > gpurun_out/r2c_check_n$N.log 2>&1; echo "rc $?" >> gpurun_out/r2c_check_n$N.log
timeout 600 $TR tests/scripts/multigpu_bands_check.py --height 3000 --width 2500 --kind 0 --sigma 5.0 > gpurun_out/r2c_check5_n$N.log 2>&1; echo "rc $?" >> gpurun_out/r2c_check5_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err; echo "rc $?" >> gpurun_out/r2c_bench_n$N.err
grep -h '"check"' gpurun_out/r2c_check_n$N.log gpurun_out/r2c_check5_n$N.log | cut -c1-300
tail -3 gpurun_out/r2c_check_n$N.log; tail -5 gpurun_out/r2c_bench_n$N.err
