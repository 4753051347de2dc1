# round 2, 8 GPUs, final build: BASELINE configs[3] (4K x 4096 spp, tile + spp) and configs[4] (120-frame flythrough); the 1080p scaling lines are the driver's SCALE run
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2m8f; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --no-cpu-baseline --res 3840x2160 --spp 4096 --steps 2 --warmup 3 --partition tile+spp --tile-groups 2 > $O/c4_tilespp.json 2> $O/c4_tilespp.err
echo "c4 rc=$? $(cut -c1-200 $O/c4_tilespp.json)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 -m digital_earth_b200.render \
  --config "digital-earth_b200/assets/configs/config - florida.txt" --res 1920x1080 --spp 256 --orbit 120 --textures synthetic:8192x4096 --no-frames --out-dir $O/fly > $O/c5_flythrough.log 2>&1
echo "c5 rc=$?"; grep -E "flythrough|rank 0" $O/c5_flythrough.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 8 --no-cpu-baseline --steps 3 --warmup 3 > $O/apollo_spp.json 2> $O/apollo_spp.err
echo "apollo rc=$? $(cut -c1-200 $O/apollo_spp.json)"
