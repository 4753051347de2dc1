cd $GRAFT_REPO_ROOT
for v in "" _o1 _o2 _o3; do
  echo "=== variant libde$v"
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --spp 16 --modes wavefront 2>&1 | grep -v "^scene"
done > gpurun_out/sweep5.log 2>&1
timeout 900 python -m pytest tests/test_gpu_render.py -x -q -k "not converges" > gpurun_out/sweep5_tests.log 2>&1
tail -40 gpurun_out/sweep5.log; tail -5 gpurun_out/sweep5_tests.log
