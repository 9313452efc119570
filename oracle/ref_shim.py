"""
ref_shim.py -- run the UNMODIFIED reference sources (Python 2.7) under Python 3 / numpy 2.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.

The reference (/root/reference, read-only, Python 2.7) cannot be imported as is.  This shim reads the
reference *.py files from where they lie, applies a fixed list of mechanical, semantics-preserving
source patches IN MEMORY (print statements, has_key, xrange, integer-division sites, bytes literals,
numpy-2 removals; the list mirrors SURVEY.md Appendix D) and exec's the result into private module
objects.  No reference source is copied into this repository; only the patch list lives here.

It is used for two things, both only inside this build container (the GPU box has no /root/reference):
  * oracle/make_golden.py drives the shimmed reference to produce the committed fixtures under
    tests/golden/ (the pinning of the oracle restatement), and
  * tests/test_oracle_vs_reference.py (skipped when /root/reference is absent) compares the
    restatement in oracle/mrc_oracle with the shimmed reference live.

Every patch carries the number of sites it must hit; a mismatch raises, so a changed reference
cannot be silently mis-patched.
"""
import os
import re
import sys
import types

REF_DIR = os.environ.get("MRC_REFERENCE_DIR", "/root/reference")

# module load order (dependencies first)
_MODULES = ["audiofile", "window", "quantize", "mdct", "bitalloc", "psychoac", "ms_stereo",
            "bitpack", "huffman", "codecThem", "pcmfile", "pacfileThem"]

_MAIN_RE = re.compile(r'^if __name__\s*==\s*"__main__"\s*:', re.M)


def _sub(src, pat, rep, count, fname, flags=0):
    new, n = re.subn(pat, rep, src, flags=flags)
    if count is not None and n != count:
        raise RuntimeError("ref_shim: patch %r hit %d sites in %s, expected %d" % (pat, n, fname, count))
    return new


def _patch(name, src):
    f = name + ".py"
    src = src.replace("\r\n", "\n").replace("﻿", "")
    # drop the __main__ self-test blocks (py2 print statements, missing modules)
    m = _MAIN_RE.search(src)
    if m:
        src = src[:m.start()]
    # ---- generic py2 -> py3 ----
    src = _sub(src, r'^(\s*)print (.*)$', r'\1print(\2)', None, f, re.M)
    src = _sub(src, r'(\w+)\.has_key\(([^()]+)\)', r'(\2 in \1)', None, f)
    src = src.replace("xrange", "range")
    src = src.replace(".tostring()", ".tobytes()")
    src = src.replace("np.fromstring", "np.frombuffer")
    src = _sub(src, r'\bnp\.float\b', 'np.float64', None, f)
    src = src.replace("import Queue as queue", "import queue")

    if name == "mdct":
        # every N/2 used as a size / range bound / slice index is py2 integer division
        src = _sub(src, r'N/2\b(?!\.)', 'N//2', None, f)
    elif name == "window":
        src = _sub(src, r'np\.linspace\(0,M,M\+1\)', 'np.linspace(0,M,int(M+1))', 1, f)
        src = _sub(src, r'np\.linspace\(0,\(N/2\.0\) - 1, \(N/2\.0\)\)',
                   'np.linspace(0,(N/2.0) - 1, int(N/2.0))', 1, f)
        src = _sub(src, r'np\.linspace\(\(N/2\.0\),N-1,\(N/2\.0\)\)',
                   'np.linspace((N/2.0),N-1,int(N/2.0))', 1, f)
    elif name == "quantize":
        # numpy 2: shift counts must be integers of a compatible kind
        src = _sub(src, r'np\.right_shift\(np\.uint64\(magVec\), shiftNum\)',
                   'np.right_shift(np.uint64(magVec), np.uint64(shiftNum))', 1, f)
        src = _sub(src, r'np\.left_shift\(np\.uint64\(mantMagVec\), shiftNum\)',
                   'np.left_shift(np.uint64(mantMagVec), np.uint64(shiftNum))', 1, f)
    elif name == "psychoac":
        src = _sub(src, r'range\(2,N/2-100\)', 'range(2,N//2-100)', 1, f)
        src = _sub(src, r'\(sampleRate/N\)', '(sampleRate//N)', 1, f)       # py2 int/int (Q2)
    elif name == "bitpack":
        # numpy 2 (NEP 50): python_int & np.uint8 stays uint8 and overflows on <<
        src = _sub(src, r'dataMask &= self\.data\[self\.iByte\]', 'dataMask &= int(self.data[self.iByte])', 2, f)
        src = _sub(src, r'dataMask = self\.data\[self\.iByte\]', 'dataMask = int(self.data[self.iByte])', 1, f)
        src = _sub(src, r'infoMask &= info\b', 'infoMask &= int(info)', 3, f)
    elif name == "codecThem":
        src = _sub(src, r'halfN = \(codingParams\.a \+ codingParams\.b\)/2\.', 'halfN = (codingParams.a + codingParams.b)//2', 4, f)
        src = _sub(src, r'freq\[0:N/2\]', 'freq[0:N//2]', 1, f)
        src = _sub(src, r'\(codingParams\.sampleRate\)/N\)', '(codingParams.sampleRate)//N)', 1, f)
        src = _sub(src, r"glob\(os\.path\.join\(x\[0\], '\*table\.pkl'\)\)\]",
                   "glob(os.path.join(x[0], '*table.pkl'))]; pickles = sorted(pickles)", 1, f)
        src = _sub(src, r'pickle\.load\(dill_pkl\)', "pickle.load(dill_pkl, encoding='latin1')", 2, f)
    elif name == "pcmfile":
        for s in ("RIFF", "WAVE", "fmt ", "data"):
            src = src.replace('"%s"' % s, 'b"%s"' % s)
        src = src.replace('"\\0"', 'b"\\0"')
        src = _sub(src, r'\(codingParams\.bitsPerSample/BYTESIZE\)', '(codingParams.bitsPerSample//BYTESIZE)', None, f)
        src = _sub(src, r'numSamples /= nChannels', 'numSamples //= nChannels', 1, f)
    elif name == "pacfileThem":
        src = _sub(src, r"tag='PAC '", "tag=b'PAC '", 1, f)
        src = _sub(src, r'\(codingParams\.a\+codingParams\.b\)/2,', '(codingParams.a+codingParams.b)//2,', 8, f)
        src = _sub(src, r'if nBytes%BYTESIZE==0:  nBytes /= BYTESIZE', 'if nBytes%BYTESIZE==0:  nBytes //= BYTESIZE', 2, f)
        src = _sub(src, r'else: nBytes = nBytes/BYTESIZE \+ 1', 'else: nBytes = nBytes//BYTESIZE + 1', 2, f)
        src = _sub(src, r'codingParams\.a/codingParams\.nMDCTLines', 'codingParams.a//codingParams.nMDCTLines', 2, f)
        src = _sub(src, r'codingParams\.b/codingParams\.nMDCTLines', 'codingParams.b//codingParams.nMDCTLines', 2, f)
        src = _sub(src, r'codingParams\.nSamplesPerBlock/codingParams\.nSamplesShort', 'codingParams.nSamplesPerBlock//codingParams.nSamplesShort', None, f)
        # canonical (alphabetical) table order, Q3
        src = _sub(src, r"(pickles = \[y for x in os\.walk\(file_path\) for y in glob\(os\.path\.join\(x\[0\], '\*(?:tree|table)\.pkl'\)\)\])",
                   r"\1; pickles = sorted(pickles)", 4, f)
        src = _sub(src, r"(all_rev_tables = \[y for x in os\.walk\(file_path\) for y in glob\(os\.path\.join\(x\[0\], '\*_table\.revpkl'\)\)\])",
                   r"\1; all_rev_tables = sorted(all_rev_tables)", 3, f)
        src = _sub(src, r'pickle\.load\(dill_pkl\)', "pickle.load(dill_pkl, encoding='latin1')", 7, f)
        src = _sub(src, r'open\(all_rev_tables\[huffTable\]\)', "open(all_rev_tables[huffTable], 'rb')", 3, f)
        src = _sub(src, r'pickle\.load\(rev_pkl\)', "pickle.load(rev_pkl, encoding='latin1')", 3, f)
    return src


_loaded = None


def load():
    """Return a dict name -> module of the shimmed reference.  Module names are registered in
    sys.modules under a private prefix AND (temporarily, during exec) under their bare names so the
    reference's own `from window import *` lines resolve to the shimmed copies."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_DIR):
        raise FileNotFoundError("reference tree not present at %s" % REF_DIR)
    saved = {n: sys.modules.get(n) for n in _MODULES}
    mods = {}
    try:
        for name in _MODULES:
            path = os.path.join(REF_DIR, name + ".py")
            with open(path, "r", encoding="utf-8") as fh:
                src = fh.read()
            code = compile(_patch(name, src), path, "exec")
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod
            exec(code, mod.__dict__)
            mods[name] = mod
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    # the tree pickles reference class huffman.HuffmanNode by module name
    sys.modules.setdefault("huffman", mods["huffman"])
    _loaded = mods
    return mods


class in_reference_cwd(object):
    """The reference locates its Huffman tables relative to the cwd ('./training_data/')."""
    def __enter__(self):
        self._old = os.getcwd()
        os.chdir(REF_DIR)

    def __exit__(self, *a):
        os.chdir(self._old)


def available():
    return os.path.isdir(REF_DIR)
