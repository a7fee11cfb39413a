mkdir -p gpurun_out; : > gpurun_out/bands.log
timeout 600 python -m pytest tests/test_gpu_bands.py -m gpu -x -q > gpurun_out/pytest_tb.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tb.log
tail -2 gpurun_out/pytest_tb.log
timeout 300 python tools/run_bands.py --gens 800 >> gpurun_out/bands.log 2>&1
for ex in p2p fused dist; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/run_bands.py --gens 800 --k 8 --exchange $ex >> gpurun_out/bands.log 2>&1
done
grep -E '^\{|rror' gpurun_out/bands.log | tail
