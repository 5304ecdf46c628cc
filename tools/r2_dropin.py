"""Developer probe: the reference's alignment driver on the replacement library (bench.py's e2e_dropin), with the
coalescer's batch statistics."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench, unicycler_b200 as ub
ub.load_library()
d, jobs, reads = bench.load_workload()
cells = sum(ub.chain_cells(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])[0] for j in jobs)
r = bench.dropin_e2e(ub, cells, int(sys.argv[1]) if len(sys.argv) > 1 else 8)
print('DROPIN threads %s window %s: %.1f ms per 30 reads (%.1f GCUPS), python alone %.1f ms' %
      (r['python_threads'], os.environ.get('UNICYCLER_B200_COALESCE_US', 'default'), r['ms_per_step'], r['value'], r['ms_python_only']))
