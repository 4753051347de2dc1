# round 2, GPU call 12 (last 2.9 GPU-minutes): clear-tiles-last work order (option tile_order): identity tests, then the 128-spp launch with / without
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
timeout 70 python -m pytest tests/test_gpu_render.py -q -x -k "same_paths or space_tile or tile_partition" > $O/pytest_tile_order.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_tile_order.log | cut -c1-200
timeout 60 python tools/launch_curve.py --scenes "Apollo 11" --spps 8,128 --reps 2 --timeline-spps 128 --variants "space_tiles=1,space_async=1,tile_order=0;space_tiles=1,space_async=1,tile_order=1" > $O/launch_curve_tile_order.log 2>&1
cat $O/launch_curve_tile_order.log | cut -c1-260
