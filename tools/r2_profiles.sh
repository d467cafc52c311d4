# Round-2 evidence run on one B200: tests, bench lines, ncu launch list + full captures, config matrix.  Outputs in gpurun_out/r02_*.
set -x
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "rc $?" >> $O/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err
# every launch of two steps of an 18-frame batch (two 9-frame chunks per step) with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --frames 18 --steps 2 --warmup 3 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_launches.log 2>&1
# the dominant kernel: one 9-frame launch (the batch pipeline's launch shape) and one 4-frame launch (round 1's capture shape)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o $O/r02_front3_9f python bench.py --frames 9 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_front9.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o $O/r02_front3_4f python bench.py --frames 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_front4.log 2>&1
# the list-driven hysteresis kernels as they run in the pipeline: caches NOT flushed between replays (what they read was just written
# by the front kernel of the same chunk), three chunks in flight
timeout 600 ncu --set full --cache-control none --clock-control none -k regex:ccl_sparse -s 6 -c 6 -f -o $O/r02_hyst_inpipe python bench.py --frames 27 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_hyst.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ccl_sparse -s 6 -c 2 -f -o $O/r02_hyst_cold python bench.py --frames 27 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_hyst_cold.log 2>&1
timeout 1500 python tests/scripts/config_matrix.py --out $O/r02_config_matrix.json > $O/r02_config_matrix.log 2>&1
tail -3 $O/r02_pytest_gpu.log; tail -2 $O/r02_bench_n1.err; ls -la $O | grep r02
