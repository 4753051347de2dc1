"""Static code size per stage function of the wavefront kernel (bytes of SASS, 16 B per instruction)."""
import collections, re, subprocess, sys
cubin, tag, srcpath = sys.argv[1:4]
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
size = collections.Counter(); chain, infn, run = [], False, False
for ln in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not run: chain = []
        run = True; chain.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    if ln.startswith('.text.') or ln.lstrip().startswith('.section'): infn = tag in ln
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m:
        run = False
        if infn:
            fns = [func(l[1]) for l in chain if l[0] == src_file]
            below = [f for f in fns if f != 'k_render_wavefront']
            size[below[-1] if below else (fns[-1] if fns else '?')] += 16
print('total %d B' % sum(size.values()))
for k, v in size.most_common(): print('%-28s %6d B' % (k, v))
