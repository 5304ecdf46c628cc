"""ncu target: a few batch calls on a golden set (default sample) so that the k-mer join kernels of kmerjoin.cu can be
captured with  ncu --set full -k regex:'insertKernel|fillKernel|rankKernel|probeKernel|emitKernel|segmentScanKernel'."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import load_golden
name = sys.argv[1] if len(sys.argv) > 1 else 'sample'
d = load_golden('semiglobal_%s.json.gz' % name)
h = ub.new_ref_seqs()
for n_, s_ in d['refs']:
    ub.add_ref_seq(h, n_, s_)
reads = [r for r in d['reads'] if r[0] in d['expected']]
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
print('join', ub.last_join_stats())
