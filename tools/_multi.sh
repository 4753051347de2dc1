cd $GRAFT_REPO_ROOT
O=gpurun_out/final
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tools/check_fused_resolve.py > $O/check_fused_n2.log 2>&1; echo "rc=$?" >> $O/check_fused_n2.log
grep -n "fused vs\|Error\|rc=" $O/check_fused_n2.log | tail -5
