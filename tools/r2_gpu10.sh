# round 2, GPU call 10: pool size (shading state in global memory: WF_COLD), code size (WF_SHRINK), 28 / 24 warps; speed, counters, icache counters, identity
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j; mkdir -p $O
M=sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
V='"" _shrink _cold _coldshrink _cold24 _w28'
for v in "" _shrink _cold _coldshrink _cold24 _w28; do
  echo "=== variant [$v]" >> $O/sweep.log
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 64 --modes wavefront --count >> $O/sweep.log 2>&1
  DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 ncu --metrics $M --clock-control none -k regex:k_render_wavefront -s 1 -c 1 --csv --log-file $O/icache$v.csv python tools/quick_bench.py --res 1920x1080 --tex 8192x4096 --spp 32 --modes wavefront --scenes "Apollo 11" > $O/ncu$v.log 2>&1
done
grep -E "variant|wavefront |stage share" $O/sweep.log | cut -c1-130
grep -E "stage share" $O/sweep.log | sed 's/.*|//' 
for v in "" _shrink _cold _coldshrink _cold24 _w28; do echo "[$v]"; python - $O/icache$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
for r in rows[1:]:
    print('   %-90s %s %s' % (r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')]))
PY
done
for v in _cold _coldshrink; do DE_LIB_PATH=$PWD/digital-earth_b200/libde$v.so timeout 300 python -m pytest tests/test_gpu_render.py -q -k "same_paths or space_tile or second_moment or checkpoint_resume or progressive" 2>&1 | tail -n 2; done
