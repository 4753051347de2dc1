# round 2, GPU call 11: scheduling tunables on top of the 2464-path pool (WF_COLD); ncu --set full of the new default (Apollo, 32 spp)
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2k; mkdir -p $O
for v in "" _ma28 _rf6 _rf14 _st32 _ph64 _fuse; do
  echo "=== variant [$v]" >> $O/sweep.log
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 64 --modes wavefront --count >> $O/sweep.log 2>&1
done
grep -E "variant|wavefront " $O/sweep.log | cut -c1-130
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_Apo_cold python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 32 --modes wavefront --scenes "Apollo 11" > $O/ncu_Apo.log 2>&1
ls -la $O
