mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
: > gpurun_out/sweep.log
timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
ENVS=16384 timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
SIDE=64 ENVS=65536 REPL=1 timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
N=32768 KS=4,8,16 timeout 300 python tools/sweep_life.py >> gpurun_out/sweep.log 2>&1
N=65536 KS=8,16 timeout 300 python tools/sweep_life.py >> gpurun_out/sweep.log 2>&1
N=65536 KS=8 CGL_TB_ROWS=512 timeout 300 python tools/sweep_life.py >> gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
