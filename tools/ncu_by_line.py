"""Join an ncu source-page CSV (per SASS address) with nvdisasm line info of the same cubin and
aggregate executed warp-instructions / active threads / stall samples per source region."""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kernel_tag = sys.argv[1], sys.argv[2], sys.argv[3]
src_file = sys.argv[4] if len(sys.argv) > 4 else 'de_wavefront.cu'
INNER = len(sys.argv) > 6 and sys.argv[6] == 'inner'
sass = subprocess.run(['nvdisasm', '-gi', '-c', cubin], capture_output=True, text=True).stdout
addr2loc, chain, infn, marker_run = {}, [], False, False
for ln in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not marker_run:
            chain = []
        marker_run = True
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    if ln.startswith('.text.') or ln.lstrip().startswith('.section'):
        infn = kernel_tag in ln
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m:
        marker_run = False
        if infn and chain:
            addr2loc[int(m.group(1), 16)] = list(chain)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, it, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    loc = addr2loc.get(a - base)
    inst, thr, smp = int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)
    key = ('?', 0)
    if loc:
        wf = [l for l in loc if l[0] == src_file]
        key = (wf[0] if INNER else wf[-1]) if wf else loc[-1]   # innermost / outermost line in the kernel's own file
    for t, v in zip((0, 1, 2), (inst, thr, smp)):
        agg[key][t] += v
        tot[t] += v
print('total warp-inst %.3e  avg threads %.2f  samples %d' % (tot[0], tot[1] / max(tot[0], 1), tot[2]))
print('%-28s %8s %7s %8s' % ('outermost line', 'inst %', 'thr/in', 'stall %'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    print('%-28s %7.2f%% %7.2f %7.2f%%' % ('%s:%d' % k, 100.0 * v[0] / tot[0], v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1)))
