cd $GRAFT_REPO_ROOT
O=gpurun_out/final3
mkdir -p $O
timeout 400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -4 $O/pytest_gpu.log; cut -c1-260 $O/bench_n1.json
