mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err
cat gpurun_out/bench.log gpurun_out/bench_ref.log; tail -5 gpurun_out/bench.err
