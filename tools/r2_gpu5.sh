# round 2, GPU call 5: final-build evidence on one GPU -- bench (both arms), smoke, ncu launch list + full captures + DRAM traffic, sanitizer
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python bench.py > $O/bench_apollo_n1.json 2> $O/bench_apollo_n1.err; echo "bench rc=$?"
cut -c1-600 $O/bench_apollo_n1.json
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; cut -c1-300 $O/bench_reference_arm.json
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for sc in "Apollo 11" "florida" "sunset hurricane"; do
  tag=$(echo $sc | cut -c1-3)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -f -o $O/wf_$tag python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 8 --modes wavefront --scenes "$sc" > $O/ncu_$tag.log 2>&1
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_render_wavefront -s 1 -c 1 --csv --log-file $O/traffic_1024spp.csv python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 1024 --modes wavefront --scenes "Apollo 11" > $O/traffic.log 2>&1
timeout 600 python bench.py --scene florida --res 640x360 --spp 64 --tex 2048x1024 --cpu-res 640x360 --cpu-spp 64 > $O/bench_florida_c1.json 2> $O/bench_florida_c1.err; cut -c1-300 $O/bench_florida_c1.json
timeout 900 python bench.py --scene sunset --spp 2048 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_sunset_c3.json 2> $O/bench_sunset_c3.err; cut -c1-300 $O/bench_sunset_c3.json
bash tools/r2_sanitize.sh
ls -la $O
