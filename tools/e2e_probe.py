"""End-to-end probe: wall time per semiGlobalAlignment batch call on sample_data, with and without the nvidia-smi sampler."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
import bench as B
d, jobs, reads = B.load_workload()
h = ub.new_ref_seqs()
for name, seq in d['refs']:
    ub.add_ref_seq(h, name, seq)
names = [r[0] for r in reads]; seqs = [r[1] for r in reads]; hits = [r[2] for r in reads]
def loop(tag, n=6):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        ub.semi_global_alignment_batch(names, seqs, hits, h, B.SCHEME, 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, ' '.join('%.1f' % t for t in ts), flush=True)
loop('plain')
s = B.ClockSampler(0)
loop('sampler')
print(s.stop())
loop('plain2')
if len(sys.argv) > 1:
    import torch
    torch.cuda.init()
    loop('torch')
