"""The reference's own long-read alignment driver, unmodified, on the sample_data fixture (BASELINE configs[1]):

    python dropin_align_harness.py <pydist copy> <out.json> <threads> [repeats]

<pydist copy> is a scratch copy of oracle/_ref/pydist whose unicycler/cpp_functions.so is either the reference
library or libunicycler_b200.so (with UNICYCLER_B200_FORWARD_LIB set for minimap).  Reads and references come from
tests/golden/semiglobal_sample.json.gz (written to FASTA / FASTQ files first, which is what the driver wants).
unicycler_align.semi_global_align_long_reads (unicycler_align.py:87-236) runs minimap, then calls the per-read C ABI
from a pool of `threads` Python threads.  Output: per read, every alignment's (ref, strand, coordinates, raw and scaled
score, CIGAR) — and the wall time of the whole call and of the alignment loop alone (minimap and the low-score
calibration excluded by timing a second call of seqan_alignment over the same reads)."""
import gzip
import json
import os
import sys
import time


def main():
    root, out_path, threads = sys.argv[1], sys.argv[2], int(sys.argv[3])
    repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, root)
    import unicycler.alignment
    import unicycler.log
    import unicycler.read_ref
    import unicycler.unicycler_align as ua
    unicycler.log.logger = unicycler.log.Log(log_filename=None, stdout_verbosity_level=0)
    d = json.load(gzip.open(os.path.join(here, 'golden', 'semiglobal_sample.json.gz'), 'rt'))
    fa, fq = os.path.join(root, 'sample_ref.fasta'), os.path.join(root, 'sample_reads.fastq')
    with open(fa, 'w') as f:
        for n, s in d['refs']:
            f.write('>%s\n%s\n' % (n, s))
    with open(fq, 'w') as f:
        for n, s, _ in d['reads']:
            f.write('@%s\n%s\n+\n%s\n' % (n, s, 'I' * len(s)))
    scheme = unicycler.alignment.AlignmentScoringScheme('3,-6,-5,-2')
    times = []
    for rep in range(repeats):
        refs = unicycler.read_ref.load_references(fa)
        read_dict, read_names, _ = unicycler.read_ref.load_long_reads(fq)
        # time only the alignment loop: wrap the thread pool's worker
        t_loop = [0.0]
        orig = ua.seqan_alignment
        first = [None]

        def timed(*a, **k):
            if first[0] is None:
                first[0] = time.perf_counter()
            r = orig(*a, **k)
            t_loop[0] = time.perf_counter() - first[0]
            return r
        ua.seqan_alignment = timed
        # wall-clock interval during which at least one thread is inside the C call (union of the call intervals)
        orig_c = ua.semi_global_alignment
        spans = []

        def timed_c(*a, **k):
            t = time.perf_counter()
            r = orig_c(*a, **k)
            spans.append((t, time.perf_counter()))
            return r
        ua.semi_global_alignment = timed_c
        t0 = time.perf_counter()
        aligned = ua.semi_global_align_long_reads(refs, fa, read_dict, read_names, fq, threads, scheme, [None], False, 10, None,
                                                  None, 0, 0, None, 0)
        total = time.perf_counter() - t0
        ua.seqan_alignment = orig
        ua.semi_global_alignment = orig_c
        spans.sort()
        union, cur_a, cur_b = 0.0, None, None
        for a_, b_ in spans:
            if cur_b is None or a_ > cur_b:
                if cur_b is not None:
                    union += cur_b - cur_a
                cur_a, cur_b = a_, b_
            else:
                cur_b = max(cur_b, b_)
        if cur_b is not None:
            union += cur_b - cur_a
        times.append((total, t_loop[0], union))
    # the same loop once more with the library answering from a cache: what the reference's Python alone costs
    # (Alignment objects, tally_up_score_and_errors, SAM lines ... under the GIL)
    cache = {}
    orig_c = ua.semi_global_alignment

    def fill(*a, **k):
        r = orig_c(*a, **k)
        cache[a[0]] = r
        return r
    ua.semi_global_alignment = fill
    refs = unicycler.read_ref.load_references(fa)
    read_dict, read_names, _ = unicycler.read_ref.load_long_reads(fq)
    ua.semi_global_align_long_reads(refs, fa, read_dict, read_names, fq, threads, scheme, [None], False, 10, None, None, 0, 0, None, 0)
    ua.semi_global_alignment = lambda *a, **k: cache[a[0]]
    refs = unicycler.read_ref.load_references(fa)
    read_dict, read_names, _ = unicycler.read_ref.load_long_reads(fq)
    first = [None]
    last = [0.0]
    orig = ua.seqan_alignment

    def timed2(*a, **k):
        if first[0] is None:
            first[0] = time.perf_counter()
        r = orig(*a, **k)
        last[0] = time.perf_counter() - first[0]
        return r
    ua.seqan_alignment = timed2
    ua.semi_global_align_long_reads(refs, fa, read_dict, read_names, fq, threads, scheme, [None], False, 10, None, None, 0, 0, None, 0)
    ua.seqan_alignment = orig
    ua.semi_global_alignment = orig_c
    python_only = last[0]
    res = {}
    for name in read_names:
        res[name] = sorted([a.ref.name, '-' if a.rev_comp else '+', a.read_start_pos, a.read_end_pos, a.ref_start_pos,
                            a.ref_end_pos, a.raw_score, '%.6f' % a.scaled_score, ''.join(a.cigar_parts)] for a in aligned[name].alignments)
    json.dump(dict(threads=threads, times=times, reads=res, python_only_s=python_only), open(out_path, 'w'))
    best = min(times, key=lambda t: t[1])
    print('reads %d, alignments kept %d, threads %d: alignment loop %.3f s (whole call %.3f s; some thread inside the C call '
          'for %.3f s of the loop); the same loop with the library answering from a cache: %.3f s' %
          (len(read_names), sum(len(v) for v in res.values()), threads, best[1], best[0], best[2], python_only))


if __name__ == '__main__':
    main()
