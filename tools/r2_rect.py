"""Developer probe: latency of ONE big unbanded rectangle (fullyGlobalAlignment, unbanded) and of a batch of them."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unicycler_b200 as ub
rng = random.Random(3)
scheme = (3, -6, -5, -2)
def mut(s, rate=0.15):
    out = []
    for c in s:
        r = rng.random()
        if r < rate / 3: out.append(rng.choice('ACGT'))
        elif r < 2 * rate / 3: continue
        elif r < rate: out.append(c); out.append(rng.choice('ACGT'))
        else: out.append(c)
    return ''.join(out)
for L, n in ((4000, 1), (8000, 1), (16000, 1), (8000, 64), (8000, 600)):
    a = [''.join(rng.choice('ACGT') for _ in range(L)) for _ in range(n)]
    b = [mut(x) for x in a]
    ub.fully_global_alignment_batch(a[:1], b[:1], scheme, False, 0)
    best = 1e9
    for rep in range(3):
        ub.fully_global_alignment_batch(a, b, scheme, False, 0)
        st = ub.last_stats()
        best = min(best, st['kernel_ms'])
    print('RECT L=%d n=%d cells=%.3g kernel_ms=%.3f GCUPS=%.1f' % (L, n, st['cells'], best, st['cells'] / best / 1e6), flush=True)
