# round 2, final-build evidence on one GPU:  gpurun --timeout 2400 -- 'bash tools/r2_final.sh'
# ncu --set full captures (3 views) -> hardware-counter json for the bench line; GPU test suite; bench (both arms); smoke; launch list; DRAM traffic;
# BASELINE configs[0], [2] and (N = 1) [3]; compute-sanitizer
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2z; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
i=0
for sc in "Apollo 11" "florida" "sunset hurricane"; do
  tag=$(echo $sc | cut -c1-3); spp=$(echo 32 16 8 | cut -d' ' -f$((i+1))); i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_$tag python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp $spp --modes wavefront --scenes "$sc" > $O/ncu_$tag.log 2>&1
done
python tools/ncu_metrics_json.py $O/r2_ncu_metrics.json apollo=$O/wf_Apo.ncu-rep florida=$O/wf_flo.ncu-rep sunset=$O/wf_sun.ncu-rep > /dev/null 2>&1 && cp $O/r2_ncu_metrics.json profiles/r2_ncu_metrics.json
timeout 600 python bench.py > $O/bench_apollo_n1.json 2> $O/bench_apollo_n1.err; echo "bench rc=$?"; cut -c1-400 $O/bench_apollo_n1.json
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; cut -c1-300 $O/bench_reference_arm.json
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_render_wavefront -s 1 -c 1 --csv --log-file $O/traffic_1024spp.csv python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 1024 --modes wavefront --scenes "Apollo 11" > $O/traffic.log 2>&1
timeout 600 python bench.py --scene florida --res 640x360 --spp 64 --tex 2048x1024 --cpu-res 640x360 --cpu-spp 64 > $O/bench_florida_c1.json 2> $O/bench_florida_c1.err; cut -c1-300 $O/bench_florida_c1.json
timeout 900 python bench.py --scene sunset --spp 2048 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_sunset_c3.json 2> $O/bench_sunset_c3.err; cut -c1-300 $O/bench_sunset_c3.json
timeout 600 python bench.py --res 3840x2160 --spp 4096 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_apollo_4k_n1.json 2> $O/bench_apollo_4k_n1.err; cut -c1-300 $O/bench_apollo_4k_n1.json
timeout 300 python tools/launch_curve.py --scenes "Apollo 11" --spps 1,8,32,128 --timeline-spps 1 --variants "space_tiles=1,space_async=1" > $O/launch_curve.log 2>&1; tail -n 12 $O/launch_curve.log | cut -c1-200
timeout 1500 python -m pytest tests -q -s -m gpu --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -n 14 $O/pytest_gpu.log | cut -c1-200
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_render.py "Apollo 11" "florida" > $O/sanitizer_$tool.log 2>&1; echo "exit=$?" >> $O/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit=|accum sum" $O/sanitizer_$tool.log | tail -5
done
ls -la $O
