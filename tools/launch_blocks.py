"""Development aid: per-dense-block time table from an ncu launch list of one DenseNet-121 training step."""
import csv, collections, re, sys
lines = open(sys.argv[1]).read().splitlines()
i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
idx = [n for n, r in enumerate(rows) if 'loader_kernel' in r['Kernel Name']]
rows = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
L = [6, 12, 24, 16]
def name(r):
    n = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('rxb::', '')
    m = re.match(r'conv_gemm_kernel<(\d+), (\d), (\d)>', n)          # <BK, PROLOGUE, EPI> -> the older two-argument names
    return 'conv_gemm_kernel<%s, %s>' % (m.group(1), m.group(2)) if m else n
def us(r): return float(r['Metric Value']) / 1e3
tab = collections.defaultdict(float)
# forward: conv<64,1> launches in order
f = [r for r in rows if name(r) == 'conv_gemm_kernel<64, 1>']
assert len(f) == 116, len(f)
k = 0
for b in range(4):
    for l in range(L[b]):
        tab[(b, 'fwd1x1')] += us(f[k]); tab[(b, 'fwd3x3')] += us(f[k + 1]); k += 2
# backward: walk launches after the last forward conv
last_f = max(n for n, r in enumerate(rows) if name(r) == 'conv_gemm_kernel<64, 1>')
bw = rows[last_f + 1:]
b = 3; layer = L[3]; state = 0
seq = [r for r in bw if name(r) in ('conv_wgrad_kernel', 'conv_gemm_kernel<32, 0>', 'conv_gemm_kernel<64, 0>', 'bn_bwd_apply_kernel', 'grad_fixup_kernel', 'bn_bwd_finalize_kernel', 'bn_relu_bwd_to_G_kernel<0>', 'stem_pool_bwd_kernel')]
k = 0
def take(nm):
    global k
    assert name(seq[k]) == nm, (k, name(seq[k]), nm)
    k += 1
    return us(seq[k - 1])
# head: finalize for bn5 comes first
tab[(3, 'ew')] += take('bn_bwd_finalize_kernel')
for b in (3, 2, 1, 0):
    for l in range(L[b]):
        tab[(b, 'fixup')] += take('grad_fixup_kernel')
        tab[(b, 'wg3x3')] += take('conv_wgrad_kernel')
        tab[(b, 'dg3x3')] += take('conv_gemm_kernel<32, 0>')
        tab[(b, 'final')] += take('bn_bwd_finalize_kernel')
        tab[(b, 'apply')] += take('bn_bwd_apply_kernel')
        tab[(b, 'wg1x1')] += take('conv_wgrad_kernel')
        tab[(b, 'dg1x1')] += take('conv_gemm_kernel<64, 0>')
        tab[(b, 'final')] += take('bn_bwd_finalize_kernel')
    tab[(b, 'fixup')] += take('grad_fixup_kernel')
    if b > 0:
        while name(seq[k]) == 'conv_wgrad_kernel': tab[(b, 'trans')] += take('conv_wgrad_kernel')
        tab[(b, 'trans')] += take('conv_gemm_kernel<64, 0>')
        tab[(b, 'trans')] += take('bn_relu_bwd_to_G_kernel<0>')
        tab[(b, 'trans')] += take('bn_bwd_finalize_kernel')
    else:
        tab[(b, 'stem')] += take('stem_pool_bwd_kernel')
        tab[(b, 'stem')] += take('bn_bwd_finalize_kernel')
        tab[(b, 'stem')] += take('bn_bwd_apply_kernel')
        tab[(b, 'stem')] += take('conv_wgrad_kernel')
cols = ['fwd1x1', 'fwd3x3', 'wg3x3', 'dg3x3', 'apply', 'wg1x1', 'dg1x1', 'fixup', 'final', 'trans', 'stem']
print('block ' + ' '.join('%8s' % c for c in cols) + '    total')
for b in range(4):
    print('%5d ' % b + ' '.join('%8.0f' % tab[(b, c)] for c in cols) + ' %8.0f' % sum(tab[(b, c)] for c in cols))
print('  all ' + ' '.join('%8.0f' % sum(tab[(b, c)] for b in range(4)) for c in cols))
