# round 2, GPU call 4: idle back-off variants on the final (no-band) kernel, the corrected band walk (tests + speed), full GPU test suite
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2d; mkdir -p $O
for v in "" _bo1000 _bo2000 _ie3 _ie5 _rf14bo _bands; do
  echo "=== variant [$v]" >> $O/sweep.log
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 64 --modes wavefront --count >> $O/sweep.log 2>&1
done
grep -E "variant|wavefront " $O/sweep.log | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_bounds.py -q -s -k "rmo_band" > $O/pytest_bands.log 2>&1; echo "rc=$?" >> $O/pytest_bands.log
grep -E "rmo bands|passed|failed" $O/pytest_bands.log | cut -c1-300
DE_LIB_PATH=$PWD/digital-earth_b200/libde_bands.so timeout 900 python -m pytest tests/test_gpu_render.py -q -s -k "4096spp" > $O/pytest_bands_image.log 2>&1; echo "rc=$?" >> $O/pytest_bands_image.log
grep -E "^\[|4096|passed|failed" $O/pytest_bands_image.log | cut -c1-330
timeout 1800 python -m pytest tests -q -s -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 12 $O/pytest_gpu.log | cut -c1-300
