# round 2, GPU call 7: fused first-trip stages (WF_FUSE_RMO / WF_FUSE_CLOUD) against the unfused build: identity tests, speed, counters, ncu of the fused build
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
for v in _f0 "" _fc; do
  echo "=== variant [$v]" >> $O/sweep.log
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 64 --modes wavefront --count >> $O/sweep.log 2>&1
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python -m pytest tests/test_gpu_render.py -q -k "same_paths or space_tile or second_moment" > $O/pytest_identity$v.log 2>&1; echo "variant [$v] identity rc=$?"; tail -n 3 $O/pytest_identity$v.log | cut -c1-200
done
grep -E "variant|wavefront |stage share" $O/sweep.log | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_Apo_fuse python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 8 --modes wavefront --scenes "Apollo 11" > $O/ncu_Apo.log 2>&1
ls -la $O
