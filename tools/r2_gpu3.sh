# round 2, GPU call 3: rmo altitude bands vs none, scheduling-parameter sweep (1080p, 8k textures, 32 spp, counters), band tests
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
for v in "" _nobands _bo256 _bo1000 _rf6 _rf8 _b128 _rf14; do
  echo "=== variant [$v]" >> $O/sweep.log
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 32 --modes wavefront --count >> $O/sweep.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_bounds.py -q -s -k "rmo" > $O/pytest_bands.log 2>&1; echo "rc=$?" >> $O/pytest_bands.log
timeout 900 python -m pytest tests/test_gpu_render.py -q -s -k "4096spp or c1_florida or megakernel or space_tile" > $O/pytest_image.log 2>&1; echo "rc=$?" >> $O/pytest_image.log
grep -E "variant|wavefront" $O/sweep.log | cut -c1-250; grep -E "^\[|rmo bands|passed|failed" $O/pytest_bands.log | cut -c1-300; grep -E "^\[|4096|C1|passed|failed" $O/pytest_image.log | cut -c1-330
