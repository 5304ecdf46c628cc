"""Developer probe: host stage only (UNICYCLER_B200_HOST_ONLY, no GPU needed) of the batch call on a golden set or on
synthetic 20 kb reads; prints the per-stage thread-ms of seeding."""
import os, sys, time
os.environ['UNICYCLER_B200_HOST_ONLY'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import load_golden
name = sys.argv[1] if len(sys.argv) > 1 else 'sample'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if name == 'synth':
    d = load_golden('semiglobal_synth.json.gz')
else:
    d = load_golden('semiglobal_%s.json.gz' % name)
h = ub.new_ref_seqs()
for n_, s_ in d['refs']:
    ub.add_ref_seq(h, n_, s_)
reads = [r for r in d['reads'] if r[0] in d['expected']]
args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), int(os.environ.get('SENS', '0')))
ts = []
for rep in range(reps):
    t0 = time.perf_counter(); ub.semi_global_alignment_batch(*args); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print('HOST set=%s reads=%d ms: min %.1f median %.1f' % (name, len(reads), ts[0], ts[len(ts) // 2]))
