# Two-GPU checks:  gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_validate_2gpu.sh'
# fused peer-memory resolve == ncclReduce + resolve (tools/check_fused_resolve.py), bench.py with both exchange steps.
cd $GRAFT_REPO_ROOT
O=gpurun_out/final
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tools/check_fused_resolve.py > $O/check_fused_n2.log 2>&1; echo "rc=$?" >> $O/check_fused_n2.log
timeout 900 $TR --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 900 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --exchange fused > $O/bench_n2_fused.json 2> $O/bench_n2_fused.err
grep -n "fused vs\|rc=" $O/check_fused_n2.log | tail -3; cut -c1-330 $O/bench_n2.json; cut -c1-330 $O/bench_n2_fused.json
