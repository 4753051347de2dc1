cd $GRAFT_REPO_ROOT
for v in "" _sf1 _sf3 _sf3bo256 _bo256 _w24 _w28; do
  echo "=== variant libde$v"
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --spp 16 --modes wavefront 2>&1 | grep -v "^scene"
done > gpurun_out/sweep7.log 2>&1
tail -40 gpurun_out/sweep7.log
