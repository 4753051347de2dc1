cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "preview or cli or library" > gpurun_out/sweep9_tests.log 2>&1
timeout 600 python tools/quick_bench.py --res 1920x1080 --spp 4 --modes preview,wavefront --scenes florida > gpurun_out/sweep9.log 2>&1
tail -15 gpurun_out/sweep9_tests.log; tail -5 gpurun_out/sweep9.log
