set -x
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err; echo "rc $?" >> gpurun_out/r2c_bench_n$N.err
timeout 300 $TR tests/scripts/multigpu_bands_check.py --height 4096 --width 4096 --kind 1 > gpurun_out/r2c_check_n$N.log 2>&1; echo "rc $?" >> gpurun_out/r2c_check_n$N.log
tail -3 gpurun_out/r2c_bench_n$N.err; grep -h '"bands_all"' gpurun_out/r2c_check_n$N.log
