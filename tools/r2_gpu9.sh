# round 2, GPU call 9: instruction-cache capacity probe (straight-line bodies of 4..192 KB, 32 warps/SM, all SMs) with the icc / gcc counters
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
./tools/probes/icache_probe > $O/icache_probe.txt 2>&1; cat $O/icache_probe.txt
timeout 600 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $O/icache_probe_ncu.csv ./tools/probes/icache_probe > /dev/null 2>&1
python - $O/icache_probe_ncu.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[h.index('ID')], r[h.index('Kernel Name')][:40]), {})[r[h.index('Metric Name')]] = r[h.index('Metric Value')]
for (i, k), m in d.items():
    print(i, k, 'icc hit %s%%  icc req %s  gcc req %s (%s%% of peak)  %s ns  inst %s' % (m.get('sm__icc_request_hit_rate.pct'), m.get('sm__icc_requests.sum'), m.get('gcc__cache_requests_type_instruction.sum'), m.get('gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed'), m.get('gpu__time_duration.sum'), m.get('smsp__inst_executed.sum')))
PY
