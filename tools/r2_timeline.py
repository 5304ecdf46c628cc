"""Developer probe: one launch of a fixture's chain jobs with the kernel's own timelines switched on
(UNICYCLER_B200_PROFILE / _TRACEJOB / _DBG=16 are read by the engine; the logs go to stderr).  The first call warms up,
a marker line separates the second."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden

name = sys.argv[1] if len(sys.argv) > 1 else 'sample'
d = load_golden('semiglobal_%s.json.gz' % name)
jobs = golden_chain_jobs(d)
for rep in range(2):
    sys.stderr.write('[marker] call %d\n' % rep); sys.stderr.flush()
    ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    print(rep, ub.last_stats())
