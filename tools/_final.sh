cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --scene florida --res 640x360 --spp 64 --tex 2048x1024 --no-cpu-baseline --steps 10 > $O/bench_florida_c1.json 2>> $O/bench_n1.err
timeout 900 python bench.py --scene sunset --spp 2048 --no-cpu-baseline --steps 2 > $O/bench_sunset.json 2>> $O/bench_n1.err
# launch list of the bench command (light metric, no replay)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# DRAM traffic of one full 1024-spp launch (two counters, single pass)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_render_wavefront --launch-skip 1 -c 1 --csv --log-file $O/traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/bench_under_ncu2.log 2>&1
# full-set captures of short launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o $O/wf_r1j_florida python tools/quick_bench.py --res 1920x1080 --spp 2 --modes wavefront --scenes florida > $O/ncu_florida.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o $O/wf_r1j_apollo python tools/quick_bench.py --res 1920x1080 --spp 2 --modes wavefront --scenes "Apollo 11" > $O/ncu_apollo.log 2>&1
timeout 300 python tools/quick_bench.py --res 1920x1080 --spp 16 --modes wavefront --count > $O/quick_count.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench_n1.json | cut -c1-400; cat $O/bench_ref.json | cut -c1-300
