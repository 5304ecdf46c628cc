"""Small profiling target: one launch of the DP kernel on the sample_data chain jobs (device-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden
name = sys.argv[2] if len(sys.argv) > 2 else 'sample'
d = load_golden('semiglobal_%s.json.gz' % name)
jobs = golden_chain_jobs(d)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
b = ub.ChainBench(jobs, tuple(d['scheme']), jobs[0]['band'])
ms = b.run_steps(n)
print('launches', n, 'ms per launch', ms / n, ub.transfer_bytes())
