"""Throughput of the pairwise entry points (BASELINE.json configs[2], configs[3]) on synthetic inputs:
  * bridge path scoring: n pairs of ~L bp (consensus vs candidate path, ~12 % divergence), fullyGlobalAlignment /
    pathAlignment with band 1000 through ub200_globalAlignmentBatch;
  * score calibration: getRandomSequenceAlignmentScores(L, n).
Prints pairs/s and GCUPS (reference cell definition) from the engine's own counters."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unicycler_b200 as ub

scheme = (3, -6, -5, -2)
rng = random.Random(11)


def mutate(s, rate):
    out = []
    for c in s:
        r = rng.random()
        if r < rate / 3: out.append(rng.choice('ACGT'))
        elif r < 2 * rate / 3: continue
        elif r < rate: out.append(c); out.append(rng.choice('ACGT'))
        else: out.append(c)
    return ''.join(out)


for L, n in ((2000, 500), (10000, 200), (20000, 100)):
    a = [''.join(rng.choice('ACGT') for _ in range(L)) for _ in range(n)]
    b = [mutate(x, 0.12) for x in a]
    for name, fn in (('fullyGlobalAlignment', ub.fully_global_alignment_batch), ('pathAlignment', ub.path_alignment_batch)):
        fn(a[:4], b[:4], scheme, True, 1000)
        t0 = time.time()
        out = fn(a, b, scheme, True, 1000)
        dt = time.time() - t0
        st = ub.last_stats()
        print('%-22s %4d pairs x %5d bp band 1000: %.3g cells, kernel %.1f ms (%.1f GCUPS), end to end %.1f ms (%.0f pairs/s)'
              % (name, n, L, st['cells'], st['kernel_ms'], st['cells'] / st['kernel_ms'] / 1e6, dt * 1e3, n / dt), flush=True)
os.environ['UNICYCLER_B200_SEED'] = '7'
for L, n in ((100, 25000), (1000, 2000), (5000, 200)):
    t0 = time.time()
    mean, sd = ub.get_random_sequence_alignment_mean_and_std_dev(L, n, scheme)
    dt = time.time() - t0
    cells = n * (L + 1) * (L + 1)
    print('calibration L=%d n=%d: mean %.4f sd %.4f, %.3g cells, %.1f ms end to end (%.1f GCUPS)' % (L, n, mean, sd, cells, dt * 1e3, cells / dt / 1e9), flush=True)
