# Round-2 (second session) evidence run on one B200 with the final build: tests, bench lines, ncu launch list + full captures,
# probes, config matrix.  Outputs in gpurun_out/r02_*; tools/make_profiles_r02.py turns them into profiles/r02_*.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "rc $?" >> $O/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err
# every launch of two steps of an 18-frame batch (two 9-frame chunks per step) with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --frames 18 --steps 2 --warmup 3 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_launches.log 2>&1
# the dominant kernel: one 9-frame launch (the batch pipeline's launch shape) and one 4-frame launch (round 1's capture shape)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o $O/r02_front3_9f python bench.py --frames 9 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_front9.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o $O/r02_front3_4f python bench.py --frames 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_front4.log 2>&1
# the fused-BGR variant of the same kernel (9 interleaved B,G,R frames)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front3_kernel -s 5 -c 1 -f -o $O/r02_front3_bgr_9f python tools/bgr_probe.py --frames 9 --steps 1 > $O/r02_ncu_front_bgr.log 2>&1
# the list-driven hysteresis kernels as they run in the pipeline (caches NOT flushed between replays) and cold
timeout 600 ncu --set full --cache-control none --clock-control none -k regex:ccl_sparse -s 6 -c 6 -f -o $O/r02_hyst_inpipe python bench.py --frames 27 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_hyst.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ccl_sparse -s 6 -c 2 -f -o $O/r02_hyst_cold python bench.py --frames 27 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > $O/r02_ncu_hyst_cold.log 2>&1
# probes
./tools/probes/blur_probe > $O/r02_blur_probe_raw.txt 2>&1
timeout 600 python tools/chunk_sweep.py --frames 64,512 --chunks 0,6,9 > $O/r02_chunk_sweep.txt 2>&1
timeout 300 python tools/bgr_probe.py > $O/r02_bgr_fused.json 2>> $O/r02_probe.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py > $O/r02_bgr_separate.json 2>> $O/r02_probe.err
timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --steps 20 > $O/r02_bgr8k_fused.json 2>> $O/r02_probe.err
B200_CANNY_BGR_FUSED=0 timeout 300 python tools/bgr_probe.py --frames 1 --height 8192 --width 8192 --steps 20 > $O/r02_bgr8k_separate.json 2>> $O/r02_probe.err
for v in 0 1; do B200_CANNY_PDL=$v timeout 300 python tools/pdl_probe.py > $O/r02_pdl$v.json 2>> $O/r02_probe.err; done
timeout 1500 python tests/scripts/config_matrix.py --out $O/r02_config_matrix.json > $O/r02_config_matrix.log 2>&1
tail -3 $O/r02_pytest_gpu.log; tail -2 $O/r02_bench_n1.err; ls -la $O | grep r02
