cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_render.py -x -q -k "nasa" > gpurun_out/sweep10_tests.log 2>&1
tail -15 gpurun_out/sweep10_tests.log
