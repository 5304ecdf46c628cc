"""Synthetic long-read workload (BASELINE.json configs[4], reduced): random reference, nanopore-like reads
(~15 % errors), hit strings synthesised from the ground truth.  Aligns all reads in one batch call on the GPU,
checks a subset against the unmodified reference library (oracle/_ref) and prints throughput."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import unicycler_b200 as ub
from oracle_lib import REF_LIB, mask_semi_global

ref_len = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 100
read_len = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
n_check = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rng = random.Random(1)
ref = ''.join(rng.choice('ACGT') for _ in range(ref_len))
comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def noisy(s, rng):
    out = []
    for c in s:
        r = rng.random()
        if r < 0.05:
            out.append(rng.choice('ACGT'))
        elif r < 0.10:
            continue
        elif r < 0.15:
            out.append(c); out.append(rng.choice('ACGT'))
        else:
            out.append(c)
    return ''.join(out)


rng2 = random.Random(2)
reads = []
for k in range(n_reads):
    L = int(read_len * rng2.uniform(0.9, 1.1))
    start = rng2.randint(0, ref_len - L)
    frag = ref[start:start + L]
    strand = '+' if rng2.random() < 0.5 else '-'
    if strand == '-':
        frag = ''.join(comp[c] for c in reversed(frag))
    seq = noisy(frag, rng2)
    reads.append(('read%d' % k, seq, '0,%d,%s,ref,%d,%d' % (len(seq), strand, start, start + L)))
scheme = (3, -6, -5, -2)
h = ub.new_ref_seqs()
ub.add_ref_seq(h, 'ref', ref)
t0 = time.time()
out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, scheme, 0)
dt = time.time() - t0
st = ub.last_stats()
naln = sum(len(o.split(';')) - 1 for o in out)
print('reads %d x %d bp vs %d bp reference: %d alignments, %.3g DP cells, kernel %.1f ms (%.1f GCUPS), end-to-end %.2f s (%.1f reads/s, %.1f GCUPS)'
      % (n_reads, read_len, ref_len, naln, st['cells'], st['kernel_ms'], st['cells'] / st['kernel_ms'] / 1e6, dt, n_reads / dt,
         st['cells'] / dt / 1e9), flush=True)
if os.path.isfile(REF_LIB) and n_check > 0:
    from refdriver import AbiLib
    lib = AbiLib(REF_LIB)
    hr = lib.new_refs([('ref', ref)])
    bad = 0
    t0 = time.time()
    for r, o in list(zip(reads, out))[:n_check]:
        want = lib.semi_global(r[0], r[1], r[2], hr, scheme)
        if mask_semi_global(o) != mask_semi_global(want):
            bad += 1
            print('MISMATCH', r[0], o[:150], want[:150])
    print('checked %d reads against the reference library (%.1f s on one host core): %d mismatches' % (n_check, time.time() - t0, bad))
