cd $GRAFT_REPO_ROOT
O=gpurun_out/final
mkdir -p $O/fly
for v in _nofold "" _nofold ""; do
  echo "=== variant libde$v"
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --spp 16 --modes wavefront 2>&1 | grep -v "^scene"
done > gpurun_out/sweep13.log 2>&1
timeout 900 python -m pytest tests/test_gpu_render.py -x -q -k "not converges and not nasa" > gpurun_out/sweep13_tests.log 2>&1
timeout 900 python bench.py --res 3840x2160 --spp 4096 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_4k_n1.json 2> $O/bench_4k_n1.err
START=$(date +%s.%N)
timeout 900 python -m digital_earth_b200.render --config "digital-earth_b200/assets/configs/config - florida.txt" --res 1920x1080 --spp 256 --orbit 16 --textures synthetic:8192x4096 --out-dir $O/fly > $O/fly.log 2>&1
END=$(date +%s.%N)
echo "flythrough 16 frames 1080p x 256 spp, wall (incl. 8k texture synthesis + upload + PNG writes): $(echo "$END - $START" | bc) s" > $O/fly_time.txt
python - <<'PY'
from PIL import Image
import glob, os
fs = sorted(glob.glob('gpurun_out/final/fly/frame_*.png'))
print(len(fs), 'frames')
for f in fs[:1] + fs[8:9]:
    Image.open(f).convert('RGB').save(f.replace('.png', '.jpg'), quality=88)
for f in fs: os.remove(f)
PY
tail -14 gpurun_out/sweep13.log; tail -3 gpurun_out/sweep13_tests.log; cut -c1-300 $O/bench_4k_n1.json; cat $O/fly_time.txt
