# round 2 multi-GPU evidence:  gpurun --gpus N -- 'N=<N> [LIGHT=1] bash tools/r2_multi.sh'   (LIGHT: spp + tile+spp lines only)
# BASELINE configs[1] (Apollo 1080p x 1024 spp) with every partition / exchange, configs[3] (4K x 4096 spp) with spp and tile+spp,
# and at N = 8 configs[4] (120-frame orbit, frames sharded).  Every line carries the N-GPU == 1-GPU identity check.
cd $GRAFT_REPO_ROOT
N=${N:-2}
O=gpurun_out/r2m$N; mkdir -p $O
P=29600
run() {  # name, args...
  name=$1; shift; P=$((P+1))
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(cut -c1-180 $O/$name.json)"
}
run apollo_spp --steps 3 --warmup 3
run apollo_tilespp --steps 3 --warmup 3 --partition tile+spp --tile-groups 2
if [ -z "$LIGHT" ]; then
  run apollo_tilespp_fused --steps 3 --warmup 3 --partition tile+spp --tile-groups 2 --exchange fused
  if [ "$N" != "2" ]; then run apollo_tile --steps 3 --warmup 3 --partition tile; fi
fi
run c4_tilespp --res 3840x2160 --spp 4096 --steps 2 --warmup 3 --partition tile+spp --tile-groups 2
if [ -z "$LIGHT" ]; then run c4_spp --res 3840x2160 --spp 4096 --steps 2 --warmup 3; fi
if [ "$N" = "8" ]; then
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P -m digital_earth_b200.render \
    --config "digital-earth_b200/assets/configs/config - florida.txt" --res 1920x1080 --spp 256 --orbit 120 --textures synthetic:8192x4096 --no-frames --out-dir $O/fly > $O/c5_flythrough.log 2>&1
  echo "c5 rc=$?"; grep -E "flythrough|rank 0" $O/c5_flythrough.log | cut -c1-300
fi
for f in $O/*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.3f G/s' % (j['value'] / 1e9), '%.1f ms' % j['ms_per_step'], 'e2e %.3f' % (j['e2e']['value'] / 1e9), j['config']['partition'][:60], j.get('identity'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
