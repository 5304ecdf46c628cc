"""Bridge path scoring through the reference's own Python (SURVEY.md 8(c) substitute for BASELINE configs[2]).

Runs INSIDE a staged copy of the reference package (oracle/_ref/pydist, see oracle/Makefile.ref) whose
unicycler/cpp_functions.so is either the unmodified reference library or libunicycler_b200.so:

    python dropin_bridge_harness.py <pydist copy> <out.json> [direct|prefetch]

It loads test/test_assembly_graph.gfa the way test_assembly_graph.py:28 does, walks random bridges (start segment,
a few intermediate segments, end segment), synthesises a read consensus for the bridge (path sequence with 10-15 %
errors) and calls path_finding.get_best_paths_for_seq exactly like bridge_long_read.py does.  The reference's
fully_global_alignment / path_alignment wrappers are left in place but wrapped by a recorder, so the output holds the
chosen paths with their scores AND every (s1, s2, band, entry point, result) tuple that crossed the seam — the
config-3 benchmark input.

mode "prefetch" is the batch integration of INTEGRATION.md: before get_best_paths_for_seq runs, the candidate paths
are enumerated with the reference's own all_paths and scored in ONE ub200_globalAlignmentBatch call; the wrapper then
answers the reference's per-path calls from that cache.  No reference function is modified.
"""
import ctypes
import json
import os
import random
import sys


def main():
    root, out_path = sys.argv[1], sys.argv[2]
    mode = sys.argv[3] if len(sys.argv) > 3 else 'direct'
    n_bridges = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    sys.path.insert(0, root)
    import unicycler.alignment
    import unicycler.assembly_graph
    import unicycler.cpp_wrappers as cw
    import unicycler.log
    import unicycler.path_finding as pf
    from unicycler import settings
    unicycler.log.logger = unicycler.log.Log(log_filename=None, stdout_verbosity_level=0)
    graph = unicycler.assembly_graph.AssemblyGraph(os.path.join(root, 'test', 'test_assembly_graph.gfa'), 25)
    graph.remove_all_overlaps()
    scheme = unicycler.alignment.AlignmentScoringScheme('3,-6,-5,-2')
    recorded = []

    def mask(r):
        f = r.split(',', 9)
        if len(f) == 10:
            f[8] = '0'
        return ','.join(f)

    cache = {}
    orig_global, orig_path = pf.fully_global_alignment, pf.path_alignment

    def rec_global(s1, s2, sc, banded, band):
        r = cache.get((s1, s2, banded, band))
        if r is None:
            r = orig_global(s1, s2, sc, banded, band)
        recorded.append(dict(fn='global', s1=s1, s2=s2, banded=banded, band=band, result=mask(r)))
        return r

    def rec_path(s1, s2, sc, banded, band):
        r = orig_path(s1, s2, sc, banded, band)
        recorded.append(dict(fn='path', s1=s1, s2=s2, banded=banded, band=band, result=mask(r)))
        return r

    pf.fully_global_alignment, pf.path_alignment = rec_global, rec_path

    def prefetch(start, end, target_length, consensus):
        """One batch call for all candidate paths of a bridge (the reference scores them one by one,
        path_finding.py:64-86)."""
        lib = cw.C_LIB
        if not hasattr(lib, 'ub200_globalAlignmentBatch'):
            return
        min_length = min(int(round(target_length * settings.MIN_RELATIVE_PATH_LENGTH)),
                         target_length - settings.RELATIVE_PATH_LENGTH_BUFFER_SIZE)
        max_length = max(int(round(target_length * settings.MAX_RELATIVE_PATH_LENGTH)),
                         target_length + settings.RELATIVE_PATH_LENGTH_BUFFER_SIZE)
        try:
            paths = pf.all_paths(graph, start, end, min_length, max_length)
        except pf.TooManyPaths:
            return
        seqs = sorted(set(graph.get_path_sequence(p) for p in paths))
        if not seqs:
            return
        n = len(seqs)
        a = (ctypes.c_char_p * n)(*[consensus.encode()] * n)
        b = (ctypes.c_char_p * n)(*[s.encode() for s in seqs])
        res = (ctypes.c_void_p * n)()
        lib.ub200_globalAlignmentBatch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                                   ctypes.POINTER(ctypes.c_char_p)] + [ctypes.c_int] * 5 + \
                                                  [ctypes.c_bool, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.ub200_globalAlignmentBatch(n, a, b, 0, scheme.match, scheme.mismatch, scheme.gap_open, scheme.gap_extend,
                                       True, 1000, res)
        for s, p in zip(seqs, res):
            cache[(consensus, s, True, 1000)] = cw.c_string_to_python_string(p)

    rng = random.Random(4242)
    segs = sorted(graph.segments)
    bridges = []
    attempts = 0
    while len(bridges) < n_bridges and attempts < 4000:
        attempts += 1
        start = rng.choice(segs) * rng.choice([1, -1])
        path, cur = [], start
        for _ in range(rng.randint(1, 6)):
            nxt = graph.forward_links.get(cur)
            if not nxt:
                break
            cur = nxt[int(rng.random() * len(nxt))]
            path.append(cur)
        if len(path) < 2:
            continue
        end, middle = path[-1], path[:-1]
        true_seq = graph.get_path_sequence(middle)
        if not (150 <= len(true_seq) <= 9000):
            continue
        rate = 0.10 + 0.05 * rng.random()
        cons = []
        for c in true_seq:
            r = rng.random()
            if r < rate / 3:
                cons.append('ACGT'[int(rng.random() * 4)])
            elif r < 2 * rate / 3:
                continue
            elif r < rate:
                cons.append(c)
                cons.append('ACGT'[int(rng.random() * 4)])
            else:
                cons.append(c)
        consensus = ''.join(cons)
        cache.clear()
        if mode == 'prefetch':
            prefetch(start, end, len(true_seq), consensus)
        first = len(recorded)
        result, progressive = pf.get_best_paths_for_seq(graph, start, end, len(true_seq), consensus, scheme, 90.0)
        bridges.append(dict(start=start, end=end, true_path=middle, target_length=len(true_seq),
                            progressive=progressive, calls=len(recorded) - first,
                            result=[[list(p), raw, disc, '%.6f' % scaled] for p, raw, disc, scaled in result]))
    json.dump(dict(mode=mode, bridges=bridges, recorded=recorded), open(out_path, 'w'))
    print('bridges %d, alignments recorded %d (global %d, path %d), progressive searches %d' %
          (len(bridges), len(recorded), sum(1 for r in recorded if r['fn'] == 'global'),
           sum(1 for r in recorded if r['fn'] == 'path'), sum(1 for b in bridges if b['progressive'])))


if __name__ == '__main__':
    main()
