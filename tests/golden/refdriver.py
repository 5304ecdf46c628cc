"""Helpers that drive a Unicycler-ABI library (the reference build in oracle/_ref, or the
product library) through ctypes the same way unicycler/cpp_wrappers.py and
unicycler/unicycler_align.py:188-225,370-398 do.  Test infrastructure only."""
import ctypes
import gzip
import os


def load_fasta(path):
    opener = gzip.open if path.endswith('.gz') else open
    seqs, name, parts = [], None, []
    with opener(path, 'rt') as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == '>':
                if name is not None:
                    seqs.append((name, ''.join(parts)))
                name, parts = line[1:].split()[0], []
            else:
                parts.append(line)
    if name is not None:
        seqs.append((name, ''.join(parts)))
    return seqs


def load_fastq(path):
    opener = gzip.open if path.endswith('.gz') else open
    reads = []
    with opener(path, 'rt') as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().strip()
            f.readline()
            f.readline()
            reads.append((h.strip()[1:].split()[0], s))
    return reads


class AbiLib(object):
    """ctypes bindings identical to unicycler/cpp_wrappers.py:33-175 for the hot-path symbols."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.semiGlobalAlignment.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_double, ctypes.c_bool, ctypes.c_int]
        L.semiGlobalAlignment.restype = ctypes.c_void_p
        for n in ('fullyGlobalAlignment', 'pathAlignment'):
            f = getattr(L, n)
            f.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.c_int, ctypes.c_bool, ctypes.c_int]
            f.restype = ctypes.c_void_p
        L.getRandomSequenceAlignmentScores.argtypes = [ctypes.c_int] * 6
        L.getRandomSequenceAlignmentScores.restype = ctypes.c_void_p
        L.newRefSeqs.argtypes = []
        L.newRefSeqs.restype = ctypes.c_void_p
        L.addRefSeq.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.addRefSeq.restype = None
        L.deleteRefSeqs.argtypes = [ctypes.c_void_p]
        L.deleteRefSeqs.restype = None
        L.freeCString.argtypes = [ctypes.c_void_p]
        L.freeCString.restype = None
        L.semiGlobalAlignmentExhaustive.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4
        L.semiGlobalAlignmentExhaustive.restype = ctypes.c_void_p
        for n in ('startAlignment', 'endAlignment'):
            f = getattr(L, n)
            f.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4
            f.restype = ctypes.c_int
        L.overlapAlignment.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 5
        L.overlapAlignment.restype = ctypes.c_void_p
        if hasattr(L, 'minimapAlignReads'):
            L.minimapAlignReads.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int]
            L.minimapAlignReads.restype = ctypes.c_void_p

    def _str(self, ptr):
        s = ctypes.cast(ptr, ctypes.c_char_p).value.decode()
        self.lib.freeCString(ptr)
        return s

    def fully_global(self, s1, s2, scheme, banded, band):
        m, mm, go, ge = scheme
        return self._str(self.lib.fullyGlobalAlignment(s1.encode(), s2.encode(), m, mm, go, ge, banded, band))

    def path(self, s1, s2, scheme, banded, band):
        m, mm, go, ge = scheme
        return self._str(self.lib.pathAlignment(s1.encode(), s2.encode(), m, mm, go, ge, banded, band))

    def exhaustive(self, s1, s2, scheme):
        m, mm, go, ge = scheme
        return self._str(self.lib.semiGlobalAlignmentExhaustive(s1.encode(), s2.encode(), m, mm, go, ge))

    def start(self, s1, s2, scheme):
        m, mm, go, ge = scheme
        return self.lib.startAlignment(s1.encode(), s2.encode(), m, mm, go, ge)

    def end(self, s1, s2, scheme):
        m, mm, go, ge = scheme
        return self.lib.endAlignment(s1.encode(), s2.encode(), m, mm, go, ge)

    def overlap(self, s1, s2, scheme, guess):
        m, mm, go, ge = scheme
        return self._str(self.lib.overlapAlignment(s1.encode(), s2.encode(), m, mm, go, ge, guess))

    def random_scores(self, length, n, scheme):
        m, mm, go, ge = scheme
        return self._str(self.lib.getRandomSequenceAlignmentScores(length, n, m, mm, go, ge))

    def new_refs(self, refs):
        h = self.lib.newRefSeqs()
        for name, seq in refs:
            self.lib.addRefSeq(h, name.encode(), seq.encode())
        return h

    def delete_refs(self, h):
        self.lib.deleteRefSeqs(h)

    def semi_global(self, read_name, read_seq, hits, refs_handle, scheme, sensitivity=0, verbosity=0):
        m, mm, go, ge = scheme
        return self._str(self.lib.semiGlobalAlignment(read_name.encode(), read_seq.encode(), verbosity,
                                                      hits.encode(), refs_handle, m, mm, go, ge, 0.0, False,
                                                      sensitivity))

    def minimap_hits(self, ref_fasta, reads_fastq, threads=1):
        """PAF text -> {read name: 'rs,re,strand,ref,fs,fe;...'} (minimap_alignment.py:33-75)."""
        paf = self._str(self.lib.minimapAlignReads(ref_fasta.encode(), reads_fastq.encode(), threads, 0, 0))
        hits = {}
        for line in paf.splitlines():
            p = line.strip().split('\t')
            if len(p) < 12:
                continue
            hits.setdefault(p[0], []).append(','.join([p[2], p[3], p[4], p[5].split()[0], p[7], p[8]]))
        return {k: ';'.join(v) for k, v in hits.items()}


def mask_ms(result):
    """Blank field 8 (wall-clock milliseconds, scoredalignment.cpp:135) of one alignment string."""
    if not result:
        return result
    x = result.split(',', 9)
    if len(x) < 10:
        return result
    x[8] = '0'
    return ','.join(x)


def mask_semi_global(output):
    """Mask every alignment of a semiGlobalAlignment return value; keep the console field."""
    parts = output.split(';')
    return ';'.join([mask_ms(p) for p in parts[:-1]] + [parts[-1]])
