"""Developer probe: host stage only (UNICYCLER_B200_HOST_ONLY, no GPU needed) on the config-5-style synthetic 20 kb reads
of tests/golden/semiglobal_synth5.json.gz; prints the per-stage thread-ms of seeding."""
import os, sys, time
os.environ['UNICYCLER_B200_HOST_ONLY'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import unicycler_b200 as ub
from oracle_lib import load_golden
from make_golden import synth5_inputs
d = load_golden('semiglobal_synth5.json.gz')
ref, reads = synth5_inputs(d['ref_len'], d['n_reads'], d['read_len'], d['seed'])
h = ub.new_ref_seqs(); ub.add_ref_seq(h, 'ref', ref)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter()
    ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), d['sensitivity'])
    print('HOST synth5 reads=%d ms %.1f' % (len(reads), (time.perf_counter() - t0) * 1e3))
