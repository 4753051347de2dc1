cd $GRAFT_REPO_ROOT
for v in _base "" _base ""; do
  echo "=== variant libde$v"
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --spp 16 --modes wavefront 2>&1 | grep -v "^scene"
done > gpurun_out/sweep11.log 2>&1
timeout 900 python -m pytest tests/test_gpu_render.py -x -q -k "not converges and not nasa" > gpurun_out/sweep11_tests.log 2>&1
tail -40 gpurun_out/sweep11.log; tail -3 gpurun_out/sweep11_tests.log
