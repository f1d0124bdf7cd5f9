"""Seeded random FASTA/FASTQ text with the quirks the parser must survive."""
import numpy as np


def rand_seq(rng, n, p_n=0.0, p_lower=0.0, p_iupac=0.0):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).copy()
    if p_lower:
        m = rng.random(n) < p_lower
        s[m] |= 0x20
    if p_n:
        # runs of N
        starts = np.nonzero(rng.random(n) < p_n)[0]
        for st in starts:
            ln = int(rng.integers(1, 6))
            s[st:st + ln] = ord("N") if rng.random() < 0.8 else ord("n")
    if p_iupac:
        m = rng.random(n) < p_iupac
        s[m] = rng.choice(np.frombuffer(b"RYKMSWBDHV-*.", dtype=np.uint8), size=int(m.sum()))
    return s.tobytes()


def fasta(rng, n_records=3, min_len=0, max_len=400, width=60, crlf=False, blank=False, final_nl=True,
          p_n=0.01, p_lower=0.05, p_iupac=0.002, junk_prefix=False, shared=None, p_shared=0.7):
    """shared: optional list of byte strings that records copy from (so genomes overlap)."""
    nl = b"\r\n" if crlf else b"\n"
    out = []
    if junk_prefix:
        out.append(b"this is not fasta ACGTACGTACGTACGTACGTACGTACGTACGTACGT" + nl + nl)
    for r in range(n_records):
        out.append(b">rec%d some description > with @ marks" % r + nl)
        if shared and rng.random() < p_shared:
            src = shared[int(rng.integers(len(shared)))]
            a = int(rng.integers(0, max(1, len(src))))
            b = int(rng.integers(a, len(src) + 1))
            s = src[a:b]
            if rng.random() < 0.5:
                comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
                s = s.translate(comp)[::-1]
        else:
            s = rand_seq(rng, int(rng.integers(min_len, max_len + 1)), p_n, p_lower, p_iupac)
        w = width if width > 0 else max(1, len(s))
        for i in range(0, len(s), w):
            out.append(s[i:i + w] + nl)
            if blank and rng.random() < 0.1:
                out.append(nl)
    txt = b"".join(out)
    if not final_nl and txt.endswith(nl):
        txt = txt[:-len(nl)]
    return txt


def fastq(rng, source: bytes, n_reads=200, read_len=50, err=0.01, crlf=False, final_nl=True):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for r in range(n_reads):
        L = min(read_len, len(source))
        a = int(rng.integers(0, len(source) - L + 1))
        s = bytearray(source[a:a + L])
        if rng.random() < 0.5:
            s = bytearray(bytes(s).translate(comp)[::-1])
        for i in np.nonzero(rng.random(L) < err)[0]:
            s[i] = b"ACGT"[int(rng.integers(4))]
        if rng.random() < 0.02 and L:
            s[int(rng.integers(L))] = ord("N")
        q = bytes(rng.choice(np.frombuffer(b"@>+IIIIIFFF#", dtype=np.uint8), size=L))
        out.append(b"@read%d/1" % r + nl + bytes(s) + nl + b"+" + nl + q + nl)
    txt = b"".join(out)
    if not final_nl and txt.endswith(nl):
        txt = txt[:-len(nl)]
    return txt
