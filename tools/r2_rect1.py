import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unicycler_b200 as ub
rng = random.Random(3)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
a = ''.join(rng.choice('ACGT') for _ in range(L)); b = ''.join(rng.choice('ACGT') for _ in range(L))
for rep in range(2):
    ub.fully_global_alignment_batch([a], [b], (3, -6, -5, -2), False, 0)
    print('kernel_ms', ub.last_stats()['kernel_ms'], flush=True)
