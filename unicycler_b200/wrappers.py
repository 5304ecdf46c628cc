"""ctypes bindings of libunicycler_b200.so mirroring unicycler/cpp_wrappers.py (reference file:line cited per
function).  Strings returned by the library are malloc()ed there and released through freeCString, exactly
like cpp_wrappers.c_string_to_python_string (cpp_wrappers.py:126-133)."""
import ctypes
import os
from ctypes import c_bool, c_char_p, c_double, c_int, c_int64, c_void_p, POINTER

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libunicycler_b200.so')
_LIB = None


def load_library():
    """Loads the shared library (cpp_wrappers.py:23-28).  Raises if it has not been built: the product has no
    pure-Python or CPU path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIB_PATH):
        raise ImportError('libunicycler_b200.so not found at %s — run `python -c "import __graft_entry__ as g; '
                          'g.build()"` or `make -C unicycler_b200/csrc`' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.semiGlobalAlignment.argtypes = [c_char_p, c_char_p, c_int, c_char_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_double, c_bool, c_int]
    L.semiGlobalAlignment.restype = c_void_p
    for name in ('fullyGlobalAlignment', 'pathAlignment'):
        f = getattr(L, name)
        f.argtypes = [c_char_p, c_char_p, c_int, c_int, c_int, c_int, c_bool, c_int]
        f.restype = c_void_p
    L.getRandomSequenceAlignmentScores.argtypes = [c_int] * 6
    L.getRandomSequenceAlignmentScores.restype = c_void_p
    L.newRefSeqs.argtypes = []
    L.newRefSeqs.restype = c_void_p
    L.addRefSeq.argtypes = [c_void_p, c_char_p, c_char_p]
    L.addRefSeq.restype = None
    L.deleteRefSeqs.argtypes = [c_void_p]
    L.deleteRefSeqs.restype = None
    L.freeCString.argtypes = [c_void_p]
    L.freeCString.restype = None
    L.ub200_globalAlignmentBatch.argtypes = [c_int, POINTER(c_char_p), POINTER(c_char_p), c_int, c_int, c_int, c_int,
                                             c_int, c_bool, c_int, POINTER(c_void_p)]
    L.ub200_globalAlignmentBatch.restype = c_int
    L.ub200_chainAlignment.argtypes = [c_char_p, c_char_p, POINTER(c_int64), c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_char_p, c_char_p, c_int]
    L.ub200_chainAlignment.restype = c_void_p
    L.ub200_chainAlignmentBatch.argtypes = [c_int, POINTER(c_char_p), POINTER(c_char_p), POINTER(c_int64),
                                            POINTER(c_int64), c_int, c_int, c_int, c_int, c_int, POINTER(c_char_p),
                                            POINTER(c_char_p), POINTER(c_int), POINTER(c_void_p)]
    L.ub200_chainAlignmentBatch.restype = c_int
    L.ub200_semiGlobalAlignmentBatch.argtypes = [c_int, POINTER(c_char_p), POINTER(c_char_p), POINTER(c_char_p),
                                                 c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]
    L.ub200_semiGlobalAlignmentBatch.restype = c_int
    L.ub200_seedChains.argtypes = [c_char_p, c_char_p, c_int]
    L.ub200_seedChains.restype = c_void_p
    L.ub200_commonKmers.argtypes = [c_char_p, c_char_p, c_int, c_int, c_int, c_int, POINTER(ctypes.c_int32), c_int64]
    L.ub200_commonKmers.restype = c_int64
    L.ub200_lastJoinStats.argtypes = [POINTER(c_double), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64),
                                      POINTER(c_int64)]
    L.ub200_lastJoinStats.restype = None
    L.ub200_lastStats.argtypes = [POINTER(c_int64), POINTER(c_double), POINTER(c_int64), POINTER(c_double),
                                  POINTER(c_double)]
    L.ub200_lastStats.restype = None
    L.ub200_chainBenchPrepare.argtypes = [c_int, POINTER(c_char_p), POINTER(c_char_p), POINTER(c_int64),
                                          POINTER(c_int64), c_int, c_int, c_int, c_int, c_int]
    L.ub200_chainBenchPrepare.restype = c_int
    L.ub200_chainBenchRun.argtypes = []
    L.ub200_chainBenchRun.restype = c_double
    L.ub200_chainBenchFinish.argtypes = [POINTER(c_void_p)]
    L.ub200_chainBenchFinish.restype = c_int
    L.ub200_chainBenchRunSteps.argtypes = [c_int]
    L.ub200_chainBenchRunSteps.restype = c_double
    L.ub200_lastTransferBytes.argtypes = [POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]
    L.ub200_lastTransferBytes.restype = None
    L.ub200_chainCells.argtypes = [c_int, c_int, POINTER(c_int64), c_int, c_int, POINTER(c_int)]
    L.ub200_chainCells.restype = c_int64
    L.ub200_chainPlan.argtypes = [c_int, c_int, POINTER(c_int64), c_int, c_int, POINTER(ctypes.c_int32), c_int]
    L.ub200_chainPlan.restype = c_int
    L.semiGlobalAlignmentExhaustive.argtypes = [c_char_p, c_char_p, c_int, c_int, c_int, c_int]
    L.semiGlobalAlignmentExhaustive.restype = c_void_p
    for name in ('startAlignment', 'endAlignment'):
        f = getattr(L, name)
        f.argtypes = [c_char_p, c_char_p, c_int, c_int, c_int, c_int]
        f.restype = c_int
    L.overlapAlignment.argtypes = [c_char_p, c_char_p, c_int, c_int, c_int, c_int, c_int]
    L.overlapAlignment.restype = c_void_p
    L.ub200_calibrationPairs.argtypes = [c_int, c_int, ctypes.c_uint, POINTER(c_void_p), POINTER(c_void_p)]
    L.ub200_calibrationPairs.restype = c_int
    L.ub200_alignmentTallies.argtypes = [c_char_p, c_char_p, c_int, c_int, c_char_p, c_int, c_int, c_int, c_int]
    L.ub200_alignmentTallies.restype = c_void_p
    L.ub200_coalescerStats.argtypes = [POINTER(c_int64), POINTER(c_int64)]
    L.ub200_coalescerStats.restype = None
    L.ub200_setDevice.argtypes = [c_int]
    L.ub200_setDevice.restype = c_int
    L.ub200_intPeakOpsPerSec.argtypes = []
    L.ub200_intPeakOpsPerSec.restype = c_double
    L.ub200_version.argtypes = []
    L.ub200_version.restype = c_char_p
    _LIB = L
    return L


def _to_str(ptr):
    s = ctypes.cast(ptr, c_char_p).value.decode()
    load_library().freeCString(ptr)
    return s


def _cstrs(items):
    arr = (c_char_p * len(items))()
    arr[:] = [x.encode() if isinstance(x, str) else x for x in items]
    return arr


# ---- reference-compatible wrappers -------------------------------------------------------------------------------

def semi_global_alignment(read_name, read_sequence, verbosity, minimap_alignments_str, ref_seqs_ptr, match_score,
                          mismatch_score, gap_open_score, gap_extend_score, low_score_threshold, keep_bad,
                          sensitivity_level):
    """cpp_wrappers.py:47-55"""
    ptr = load_library().semiGlobalAlignment(read_name.encode(), read_sequence.encode(), verbosity,
                                             minimap_alignments_str.encode(), ref_seqs_ptr, match_score,
                                             mismatch_score, gap_open_score, gap_extend_score, low_score_threshold,
                                             keep_bad, sensitivity_level)
    return _to_str(ptr)


def fully_global_alignment(sequence_1, sequence_2, scoring_scheme, use_banding, band_size):
    """cpp_wrappers.py:90-95; scoring_scheme = (match, mismatch, gap_open, gap_extend) or an object with those
    attributes (alignment.AlignmentScoringScheme)."""
    m, mm, go, ge = _scheme(scoring_scheme)
    return _to_str(load_library().fullyGlobalAlignment(sequence_1.encode(), sequence_2.encode(), m, mm, go, ge,
                                                       use_banding, band_size))


def path_alignment(partial_seq, full_seq, scoring_scheme, use_banding, band_size):
    """cpp_wrappers.py:112-117"""
    m, mm, go, ge = _scheme(scoring_scheme)
    return _to_str(load_library().pathAlignment(partial_seq.encode(), full_seq.encode(), m, mm, go, ge, use_banding,
                                                band_size))


def semi_global_alignment_exhaustive(sequence_1, sequence_2, scoring_scheme):
    """cpp_wrappers.py:69-74"""
    m, mm, go, ge = _scheme(scoring_scheme)
    return _to_str(load_library().semiGlobalAlignmentExhaustive(sequence_1.encode(), sequence_2.encode(), m, mm, go, ge))


def start_seq_alignment(s_1, s_2, scoring_scheme):
    """cpp_wrappers.py:338-341"""
    m, mm, go, ge = _scheme(scoring_scheme)
    return load_library().startAlignment(s_1.encode(), s_2.encode(), m, mm, go, ge)


def end_seq_alignment(s_1, s_2, scoring_scheme):
    """cpp_wrappers.py:354-357"""
    m, mm, go, ge = _scheme(scoring_scheme)
    return load_library().endAlignment(s_1.encode(), s_2.encode(), m, mm, go, ge)


def overlap_alignment(s_1, s_2, scoring_scheme, guess_overlap):
    """cpp_wrappers.py:195-200 -> (overlap_1, overlap_2)"""
    m, mm, go, ge = _scheme(scoring_scheme)
    r = _to_str(load_library().overlapAlignment(s_1.encode(), s_2.encode(), m, mm, go, ge, guess_overlap))
    a, b = r.split(',')
    return int(a), int(b)


def calibration_pairs(seq_length, n, seed):
    """The sequence pairs getRandomSequenceAlignmentScores(seq_length, n, ...) aligns under UNICYCLER_B200_SEED=seed."""
    a, b = (c_void_p * n)(), (c_void_p * n)()
    load_library().ub200_calibrationPairs(seq_length, n, seed, a, b)
    return [_to_str(p) for p in a], [_to_str(p) for p in b]


def alignment_tallies(read_seq_as_aligned, ref_seq, read_start_pos, ref_start_pos, cigar, scoring_scheme):
    """The tallies of Alignment.tally_up_score_and_errors (alignment.py:142-216) computed by the library: dict with
    match_count, mismatch_count, insertion_count, deletion_count, raw_score, alignment_length, percent_identity,
    scaled_score, edit_distance; None when only soft clips remain."""
    m, mm, go, ge = _scheme(scoring_scheme)
    r = _to_str(load_library().ub200_alignmentTallies(read_seq_as_aligned.encode(), ref_seq.encode(), read_start_pos,
                                                      ref_start_pos, cigar.encode(), m, mm, go, ge))
    if not r:
        return None
    f = r.split(',')
    d = dict(match_count=int(f[0]), mismatch_count=int(f[1]), insertion_count=int(f[2]), deletion_count=int(f[3]),
             raw_score=int(f[4]), alignment_length=int(f[5]), percent_identity=float(f[6]), scaled_score=float(f[7]))
    d['edit_distance'] = d['mismatch_count'] + d['insertion_count'] + d['deletion_count']
    return d


def coalescer_stats():
    b, r = c_int64(), c_int64()
    load_library().ub200_coalescerStats(ctypes.byref(b), ctypes.byref(r))
    return dict(batches=b.value, requests=r.value)


def get_random_sequence_alignment_mean_and_std_dev(seq_length, count, scoring_scheme):
    """cpp_wrappers.py:169-175"""
    m, mm, go, ge = _scheme(scoring_scheme)
    parts = _to_str(load_library().getRandomSequenceAlignmentScores(seq_length, count, m, mm, go, ge)).split(',')
    return float(parts[0]), float(parts[1])


def new_ref_seqs():
    return load_library().newRefSeqs()


def add_ref_seq(ref_seqs_ptr, name, sequence):
    load_library().addRefSeq(ref_seqs_ptr, name.encode(), sequence.encode())


def delete_ref_seqs(ref_seqs_ptr):
    load_library().deleteRefSeqs(ref_seqs_ptr)


def _scheme(s):
    if isinstance(s, (tuple, list)):
        return tuple(int(x) for x in s)
    return s.match, s.mismatch, s.gap_open, s.gap_extend


# ---- additive batch API --------------------------------------------------------------------------------------------

def _global_batch(mode, seqs_1, seqs_2, scoring_scheme, use_banding, band_size):
    m, mm, go, ge = _scheme(scoring_scheme)
    n = len(seqs_1)
    out = (c_void_p * n)()
    rc = load_library().ub200_globalAlignmentBatch(n, _cstrs(seqs_1), _cstrs(seqs_2), mode, m, mm, go, ge, use_banding,
                                                   band_size, out)
    if rc != 0:
        raise RuntimeError('ub200_globalAlignmentBatch failed: %d' % rc)
    return [_to_str(p) for p in out]


def fully_global_alignment_batch(seqs_1, seqs_2, scoring_scheme, use_banding, band_size):
    return _global_batch(0, seqs_1, seqs_2, scoring_scheme, use_banding, band_size)


def path_alignment_batch(seqs_1, seqs_2, scoring_scheme, use_banding, band_size):
    return _global_batch(1, seqs_1, seqs_2, scoring_scheme, use_banding, band_size)


def _seed_array(seeds):
    flat = [int(x) for s in seeds for x in s]
    return (c_int64 * max(1, len(flat)))(*flat)


def chain_alignment(read_seq, trimmed_ref_seq, seeds, scoring_scheme, band_size, read_name, ref_name, ref_offset):
    m, mm, go, ge = _scheme(scoring_scheme)
    return _to_str(load_library().ub200_chainAlignment(read_seq.encode(), trimmed_ref_seq.encode(), _seed_array(seeds),
                                                       len(seeds), m, mm, go, ge, band_size, read_name.encode(),
                                                       ref_name.encode(), ref_offset))


def _chain_args(jobs):
    """jobs: list of dicts with readSeq, refSeq, seeds[, readName, refName, refOffset]."""
    n = len(jobs)
    flat, offs = [], [0]
    for j in jobs:
        for s in j['seeds']:
            flat.extend(int(x) for x in s)
        offs.append(offs[-1] + len(j['seeds']))
    seeds = (c_int64 * max(1, len(flat)))(*flat)
    offsets = (c_int64 * (n + 1))(*offs)
    return n, _cstrs([j['readSeq'] for j in jobs]), _cstrs([j['refSeq'] for j in jobs]), seeds, offsets


def chain_alignment_batch(jobs, scoring_scheme, band_size):
    m, mm, go, ge = _scheme(scoring_scheme)
    n, reads, refs, seeds, offsets = _chain_args(jobs)
    names = _cstrs([j.get('readName', 'read+') for j in jobs])
    rnames = _cstrs([j.get('refName', 'ref') for j in jobs])
    roffs = (c_int * n)(*[int(j.get('refOffset', 0)) for j in jobs])
    out = (c_void_p * n)()
    rc = load_library().ub200_chainAlignmentBatch(n, reads, refs, seeds, offsets, m, mm, go, ge, band_size, names,
                                                  rnames, roffs, out)
    if rc != 0:
        raise RuntimeError('ub200_chainAlignmentBatch failed: %d' % rc)
    return [_to_str(p) for p in out]


def semi_global_alignment_batch(read_names, read_seqs, hit_strs, ref_seqs_ptr, scoring_scheme, sensitivity_level=0):
    m, mm, go, ge = _scheme(scoring_scheme)
    n = len(read_names)
    out = (c_void_p * n)()
    rc = load_library().ub200_semiGlobalAlignmentBatch(n, _cstrs(read_names), _cstrs(read_seqs), _cstrs(hit_strs),
                                                       ref_seqs_ptr, m, mm, go, ge, sensitivity_level, out)
    if rc != 0:
        raise RuntimeError('ub200_semiGlobalAlignmentBatch failed: %d' % rc)
    return [_to_str(p) for p in out]


def seed_chains(read_seq, trimmed_ref_seq, sensitivity_level=0):
    """Host seeding stage only -> list of chains, each a list of [beginH, beginV, endH, endV, lowerDiag, upperDiag]."""
    out = _to_str(load_library().ub200_seedChains(read_seq.encode(), trimmed_ref_seq.encode(), sensitivity_level))
    parts = out.split(';')
    chains = []
    for c in parts[1:1 + int(parts[0])]:
        body = c.split(':', 1)[1]
        chains.append([[int(x) for x in s.split(',')] for s in body.split('|')] if body else [])
    return chains


def common_kmers(read_seq, ref_seq, ref_start, ref_len, k, on_device):
    """Common k-mer points [(read position, window position), ...] of a read strand and the window
    [ref_start, ref_start + ref_len) of ref_seq, in the reference's order: host join (the per-read entry point's)
    or device join (the batch path's)."""
    L = load_library()
    r, f = read_seq.encode(), ref_seq.encode()
    n = L.ub200_commonKmers(r, f, ref_start, ref_len, k, 1 if on_device else 0, None, 0)
    if n < 0:
        raise RuntimeError('device k-mer join unavailable')
    buf = (ctypes.c_int32 * (2 * max(n, 1)))()
    n2 = L.ub200_commonKmers(r, f, ref_start, ref_len, k, 1 if on_device else 0, buf, n)
    assert n2 == n
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


def last_join_stats():
    """Counters of the device k-mer join inside the last semi_global_alignment_batch call."""
    ms = c_double()
    launches, points, h2d, d2h = c_int64(), c_int64(), c_int64(), c_int64()
    load_library().ub200_lastJoinStats(ctypes.byref(ms), ctypes.byref(launches), ctypes.byref(points), ctypes.byref(h2d),
                                       ctypes.byref(d2h))
    return {'kernel_ms': ms.value, 'launches': launches.value, 'points': points.value, 'h2d_bytes': h2d.value,
            'd2h_bytes': d2h.value}


def last_stats():
    cells, launches = c_int64(), c_int64()
    kms, h2d, d2h = c_double(), c_double(), c_double()
    load_library().ub200_lastStats(ctypes.byref(cells), ctypes.byref(kms), ctypes.byref(launches), ctypes.byref(h2d),
                                   ctypes.byref(d2h))
    return dict(cells=cells.value, kernel_ms=kms.value, launches=launches.value, h2d_ms=h2d.value, d2h_ms=d2h.value)


def transfer_bytes():
    h2d, d2h, tb, ctas = c_int64(), c_int64(), c_int64(), c_int()
    load_library().ub200_lastTransferBytes(ctypes.byref(h2d), ctypes.byref(d2h), ctypes.byref(tb), ctypes.byref(ctas))
    return dict(h2d_bytes=h2d.value, d2h_bytes=d2h.value, trace_bytes=tb.value, ctas=ctas.value)


def chain_cells(read_len, ref_len, seeds, band_size):
    """(reference DP cells, sub-DP count) of one banded-chain alignment; host planner only (no GPU)."""
    n = c_int()
    cells = load_library().ub200_chainCells(read_len, ref_len, _seed_array(seeds), len(seeds), band_size, ctypes.byref(n))
    return cells, n.value


def chain_plan(read_len, ref_len, seeds, band_size, cap=8192):
    """Sub-DP plan of one banded-chain alignment: list of (kind, nH, nV, banded, lo, up, h0, v0, hNext, vNext)."""
    L = load_library()
    out = (ctypes.c_int32 * (10 * cap))()
    k = L.ub200_chainPlan(read_len, ref_len, _seed_array(seeds), len(seeds), band_size, out, cap)
    if k < 0:
        return None
    return [tuple(out[10 * i:10 * i + 10]) for i in range(min(k, cap))]


def set_device(device):
    return load_library().ub200_setDevice(int(device))


def int_peak_ops_per_sec():
    return load_library().ub200_intPeakOpsPerSec()


class ChainBench(object):
    """Device-resident benchmark harness for the banded-chain path: inputs are uploaded and planned once;
    run() launches the DP kernel on the resident inputs; finish() fetches and formats the results."""

    def __init__(self, jobs, scoring_scheme, band_size):
        m, mm, go, ge = _scheme(scoring_scheme)
        self.n, reads, refs, seeds, offsets = _chain_args(jobs)
        self._keep = (reads, refs, seeds, offsets)
        rc = load_library().ub200_chainBenchPrepare(self.n, reads, refs, seeds, offsets, m, mm, go, ge, band_size)
        if rc != 0:
            raise RuntimeError('ub200_chainBenchPrepare failed: %d' % rc)

    def run(self):
        load_library().ub200_chainBenchRun()

    def run_steps(self, steps):
        """`steps` back-to-back launches; returns total CUDA-event milliseconds."""
        return load_library().ub200_chainBenchRunSteps(int(steps))

    def finish(self, want_results=True):
        out = (c_void_p * self.n)()
        load_library().ub200_chainBenchFinish(out)
        res = [_to_str(p) for p in out]
        return res if want_results else None
