# round 2, GPU call 6 (re-entry): state of HEAD on one GPU -- bench, ncu --set full of Apollo (8 spp) + per-stage, launch list, GPU test suite
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_apollo_n1.json 2> $O/bench_apollo_n1.err; echo "bench rc=$?"
cut -c1-700 $O/bench_apollo_n1.json
timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 64 --modes wavefront --count > $O/quick_count.log 2>&1
grep -E "wavefront|stage share" $O/quick_count.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_Apo python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 8 --modes wavefront --scenes "Apollo 11" > $O/ncu_Apo.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 30 $O/pytest_gpu.log | cut -c1-300
ls -la $O
