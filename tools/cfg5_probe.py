"""Developer probe: bench.py's config-5 slice alone (N reads, default 512), optionally without the CPU reference sample."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import bench
import unicycler_b200 as ub
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
print(json.dumps(bench.config5_synthetic(ub, ub.int_peak_ops_per_sec(), os.cpu_count() or 1, n)))
