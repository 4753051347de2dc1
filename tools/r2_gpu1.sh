# round 2, GPU call 1: new parity/bounds/image-gate tests, launch-cost curve + timeline, first bench with space tiles
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 600 python tools/launch_curve.py > $O/launch_curve.log 2>&1; echo "rc=$?" >> $O/launch_curve.log
timeout 400 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 1500 python -m pytest tests -q -s -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log; cat $O/launch_curve.log | tail -40; cut -c1-400 $O/bench_n1.json
