# round 2, GPU call 2: drain diagnostics (default vs WF_DRAIN_FAST=0), ncu --set full of HEAD (Apollo, florida), fixed tests
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
timeout 400 python tools/launch_curve.py --scenes "Apollo 11,sunset hurricane" --spps 1,4,16,64 --timeline-spps 1,16 --variants "space_tiles=1,space_async=1" --cta 3 > $O/curve_drainfast.log 2>&1
DE_LIB_PATH=$PWD/digital-earth_b200/libde_nodrain.so timeout 400 python tools/launch_curve.py --scenes "Apollo 11,sunset hurricane" --spps 1,4,16,64 --timeline-spps 1,16 --variants "space_tiles=1,space_async=1" --cta 3 > $O/curve_nodrain.log 2>&1
timeout 900 python -m pytest tests/test_gpu_bounds.py -q -s -k terrain > $O/pytest_terrain.log 2>&1; echo "rc=$?" >> $O/pytest_terrain.log
timeout 600 python -m pytest tests/test_gpu_render.py -q -s -k "tile or second_moment" > $O/pytest_tiles.log 2>&1; echo "rc=$?" >> $O/pytest_tiles.log
for sc in "Apollo 11" "florida"; do
  tag=$(echo $sc | cut -c1-3)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_$tag python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 2 --modes wavefront --scenes "$sc" > $O/ncu_$tag.log 2>&1
done
tail -n 30 $O/curve_drainfast.log; tail -n 30 $O/curve_nodrain.log; tail -n 4 $O/pytest_terrain.log; tail -n 4 $O/pytest_tiles.log; ls -la $O
