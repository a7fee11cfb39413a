mkdir -p gpurun_out; : > gpurun_out/sweep.log
CGL_ENV_TMA_THREADS=256 timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
CGL_ENV_IMPL=fused timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
ENVS=16384 CGL_ENV_TMA_THREADS=256 timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
ENVS=16384 CGL_ENV_IMPL=fused timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
REPL=1 CGL_ENV_IMPL=fused timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
REPL=1 timeout 120 python tools/sweep_env.py >> gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
