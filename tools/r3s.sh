set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_pytest_gpu.log; cat $O/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "rc $?" >> $O/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
tail -2 $O/r02_bench_n1.err
