"""Test-side access to the oracle (oracle/_build/liboracle.so) and, when present, the unmodified reference
library (oracle/_ref/libunicycler_ref.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this module."""
import ctypes
import gzip
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, 'oracle', '_build', 'liboracle.so')
REF_LIB = os.path.join(ROOT, 'oracle', '_ref', 'libunicycler_ref.so')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

_COMP = str.maketrans('ACGTRYSWKMBDHVN.-?*', 'TGCAYRSWMKVHDBN.-?*')


def revcomp(s):
    """string_functions.cpp:52-79 (characters outside the table are dropped)."""
    keep = set('ACGTRYSWKMBDHVN.-?*')
    return ''.join(c for c in s if c in keep).translate(_COMP)[::-1]


def mask_ms(result):
    if not result:
        return result
    x = result.split(',', 9)
    if len(x) < 10:
        return result
    x[8] = '0'
    return ','.join(x)


def mask_semi_global(output):
    parts = output.split(';')
    return ';'.join([mask_ms(p) for p in parts[:-1]] + [parts[-1]])


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), 'rt') as f:
        return json.load(f)


def golden_chain_jobs(d):
    """Expands the chain jobs of a semiglobal_<set>.json.gz fixture into self-contained dicts."""
    refs = dict(d['refs'])
    reads = {r[0]: r[1] for r in d['reads']}
    jobs = []
    for j in d['jobs']:
        seq = reads[j['read'][:-1]]
        if j['read'][-1] == '-':
            seq = revcomp(seq)
        assert len(seq) == j['readLen']
        ref = refs[j['ref']][j['refStart']:j['refStart'] + j['refLen']]
        jobs.append(dict(readSeq=seq, refSeq=ref, seeds=j['seeds'], readName=j['read'], refName=j['ref'],
                         refOffset=j['refStart'], band=j['band'], result=j['result']))
    return jobs


class Oracle(object):
    def __init__(self):
        L = ctypes.CDLL(ORACLE_LIB)
        for n in ('oracle_fullyGlobalAlignment', 'oracle_pathAlignment'):
            f = getattr(L, n)
            f.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4 + [ctypes.c_bool, ctypes.c_int]
            f.restype = ctypes.c_void_p
        L.oracle_chainAlignment.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long),
                                            ctypes.c_int] + [ctypes.c_int] * 5 + [ctypes.c_char_p, ctypes.c_char_p,
                                                                                  ctypes.c_int]
        L.oracle_chainAlignment.restype = ctypes.c_void_p
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_lastCells.restype = ctypes.c_longlong
        L.oracle_lastGrids.restype = ctypes.c_longlong
        self.L = L

    def _s(self, p):
        s = ctypes.cast(p, ctypes.c_char_p).value.decode()
        self.L.oracle_free(p)
        return s

    def fully_global(self, s1, s2, scheme, banded, band):
        m, mm, go, ge = scheme
        return self._s(self.L.oracle_fullyGlobalAlignment(s1.encode(), s2.encode(), m, mm, go, ge, banded, band))

    def path(self, s1, s2, scheme, banded, band):
        m, mm, go, ge = scheme
        return self._s(self.L.oracle_pathAlignment(s1.encode(), s2.encode(), m, mm, go, ge, banded, band))

    def chain(self, read_seq, ref_seq, seeds, scheme, band, read_name, ref_name, ref_offset):
        m, mm, go, ge = scheme
        flat = [int(x) for s in seeds for x in s]
        arr = (ctypes.c_long * max(1, len(flat)))(*flat)
        return self._s(self.L.oracle_chainAlignment(read_seq.encode(), ref_seq.encode(), arr, len(seeds), m, mm, go,
                                                    ge, band, read_name.encode(), ref_name.encode(), ref_offset))

    def last_cells(self):
        return self.L.oracle_lastCells()

    def last_grids(self):
        return self.L.oracle_lastGrids()
