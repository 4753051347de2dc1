# What produced the round evidence under profiles/ (one B200):  gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh'
# GPU tests, smoke, both bench arms, the flythrough; ncu passes are listed in profiles/r1_bench.md.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final2/fly
O=gpurun_out/final2
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python -m digital_earth_b200.render --config "digital-earth_b200/assets/configs/config - florida.txt" --res 1920x1080 --spp 256 --orbit 16 --textures synthetic:8192x4096 --out-dir $O/fly > $O/fly.log 2>&1
rm -f $O/fly/*.png
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-260 $O/bench_n1.json; cut -c1-200 $O/bench_ref.json; grep "frames of" $O/fly.log
