"""Aggregate an ncu source-page CSV per *stage* of the wavefront kernel: every SASS instruction is
attributed to the outermost function of its inline chain below the kernel body (stage_event,
burst_track, ...; out-of-line helpers such as philox_block are their own rows).
usage: ncu_by_stage.py REP CUBIN KERNEL_TAG SRC [inner]   ('inner' adds the innermost function as a sub-key)"""
import collections, csv, re, subprocess, sys
rep, cubin, tag, srcpath = sys.argv[1:5]
inner = len(sys.argv) > 5
src_file = srcpath.split('/')[-1]
starts = []
for i, l in enumerate(open(srcpath).read().split('\n'), 1):
    m = re.match(r'(?:template <[^>]*> )?(?:DE_DEV|__device__ __noinline__|__global__ void __launch_bounds__\([^)]*\)|__global__)[^(]*?(\w+)\(', l)
    if m: starts.append((i, m.group(1)))
def func(line):
    name = '?'
    for i, n in starts:
        if i <= line: name = n
    return name
sass = subprocess.run(['nvdisasm', '-gi', '-c', cubin], capture_output=True, text=True).stdout
addr2, chain, infn, run = {}, [], False, False
for ln in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not run: chain = []
        run = True; chain.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    if ln.startswith('.text.') or ln.lstrip().startswith('.section'): infn = tag in ln
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m:
        run = False
        if infn and chain: addr2[int(m.group(1), 16)] = list(chain)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]
ia, ii, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
reasons = ['stall_no_inst', 'stall_wait', 'stall_long_sb', 'stall_short_sb', 'stall_branch_resolving', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_mio', 'stall_lg', 'stall_barrier', 'stall_sleep', 'stall_membar']
ir = {r: hdr.index(r) for r in reasons if r in hdr}
agg = collections.defaultdict(collections.Counter); base = None
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = a
    loc = addr2.get(a - base)
    key = '?'
    if loc:
        fns = [func(l[1]) for l in loc if l[0] == src_file]   # innermost ... outermost
        below = [f for f in fns if f != 'k_render_wavefront']
        key = below[-1] if below else (fns[-1] if fns else loc[-1][0])
        if inner:
            leaf = func(loc[0][1]) if loc[0][0] == src_file else loc[0][0] + ':%d' % loc[0][1]
            if leaf != key: key += ' > ' + leaf
    agg[key]['inst'] += int(r[ii] or 0); agg[key]['thr'] += int(r[it] or 0)
    for n, i in ir.items(): agg[key][n] += int(r[i] or 0)
tot = collections.Counter()
for v in agg.values(): tot.update(v)
tots = sum(tot[n] for n in ir)
print('total inst %.3e lanes %.2f ; stall samples %d: ' % (tot['inst'], tot['thr'] / tot['inst'], tots) + ' '.join('%s %.1f%%' % (n[6:], 100 * tot[n] / tots) for n in ir if tot[n] > 0.01 * tots))
print('%-34s %7s %6s %7s %7s | %s' % ('stage', 'inst%', 'lanes', 'lost%', 'stall%', ' '.join('%8s' % n[6:14] for n in list(ir)[:7])))
nrows = 60 if inner else 30
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['inst'])[:nrows]:
    st = sum(v[n] for n in ir)
    lost = (32 * v['inst'] - v['thr']) / (32.0 * tot['inst'])   # share of all issue slots x lanes wasted here
    print('%-34s %6.2f%% %6.2f %6.2f%% %6.2f%% | %s' % (k[:34], 100 * v['inst'] / tot['inst'], v['thr'] / max(v['inst'], 1), 100 * lost, 100 * st / tots, ' '.join('%7.1f%%' % (100 * v[n] / max(st, 1)) for n in list(ir)[:7])))
