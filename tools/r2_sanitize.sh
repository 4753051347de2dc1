# compute-sanitizer on the wavefront kernel's shared-memory MPMC queues (SURVEY section 5 / VERDICT r1 #6): memcheck, racecheck, synccheck
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2s; mkdir -p $O
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_render.py > $O/$tool.log 2>&1; echo "exit=$?" >> $O/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit=|accum sum" $O/$tool.log | tail -6
done
