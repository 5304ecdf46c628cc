"""Developer probe: end-to-end batch call (host strings in, result strings out) on a golden set, with the host timeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import load_golden, mask_semi_global
name = sys.argv[1] if len(sys.argv) > 1 else 'sample'
d = load_golden('semiglobal_%s.json.gz' % name)
h = ub.new_ref_seqs()
for n_, s_ in d['refs']:
    ub.add_ref_seq(h, n_, s_)
reads = [r for r in d['reads'] if r[0] in d['expected']]
args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
for rep in range(3):
    ub.semi_global_alignment_batch(*args)
ts = []
for rep in range(8):
    t0 = time.perf_counter(); out = ub.semi_global_alignment_batch(*args); ts.append((time.perf_counter() - t0) * 1e3)
bad = sum(1 for r, o in zip(reads, out) if mask_semi_global(o) != d['expected'][r[0]])
ts.sort()
print('E2E set=%s reads=%d bad=%d ms: min %.1f median %.1f max %.1f  cores=%d' % (name, len(reads), bad, ts[0], ts[len(ts) // 2], ts[-1], os.cpu_count()), ub.last_stats(), flush=True)
