"""On-disk formats either side of the hot path: the `.2bit` read container and the `.ovl` record
stream.  Host logic only (numpy); byte-exact with the reference.

ref: lib/bseq.c:93-97 (header {0,254}), :114-139 (seq2bit), :257-302 (kbit_read);
     lib/ovl.c:10-29 (varint), :109-150 (encode_ovl), :150-203 (decode_ovl).
"""
from dataclasses import dataclass
import numpy as np


@dataclass
class Reads:
    """A batch of reads in the layout the C ABI takes (include/ndgpu.h: ndg_reads_t)."""
    words: np.ndarray        # uint32, concatenated 2-bit words
    word_off: np.ndarray     # uint64, n+1
    length: np.ndarray       # uint32
    ids: np.ndarray          # uint32 numeric names (strtoul of the name)
    name_key: np.ndarray     # uint64 strcmp-order-preserving key
    name_hash: np.ndarray    # uint32 X31 hash of the decimal name
    nmask: np.ndarray = None # optional uint32: 1 bit per base (ambiguous), word i of read r at nmask[word_off[r]//2 + i]

    @property
    def n(self):
        return int(self.length.shape[0])

    @property
    def n_bases(self):
        return int(self.length.sum(dtype=np.uint64))

    def slice(self, r0, r1):
        w0, w1 = int(self.word_off[r0]), int(self.word_off[r1])
        return Reads(self.words[w0:w1], (self.word_off[r0:r1 + 1] - np.uint64(w0)).astype(np.uint64),
                     self.length[r0:r1], self.ids[r0:r1], self.name_key[r0:r1], self.name_hash[r0:r1],
                     None if self.nmask is None else self.nmask[w0 // 2:w1 // 2 + 1])


def take_reads(r: "Reads", idx):
    """a new batch made of reads idx[0], idx[1], ... of r (copies the 2-bit words; an ambiguity mask is not carried)"""
    idx = np.asarray(idx, dtype=np.int64)
    w0 = r.word_off[idx].astype(np.int64)
    nw = (r.word_off[idx + 1].astype(np.int64) - w0)
    off = np.zeros(idx.shape[0] + 1, dtype=np.uint64)
    np.cumsum(nw, out=off[1:])
    if idx.shape[0] and (nw == nw[0]).all():                       # fixed-length reads: one fancy-indexed gather
        words = r.words[(w0[:, None] + np.arange(int(nw[0]), dtype=np.int64)[None, :]).reshape(-1)]
    else:
        words = np.concatenate([r.words[a:a + n] for a, n in zip(w0.tolist(), nw.tolist())]) if idx.shape[0] else np.zeros(0, dtype=np.uint32)
    return Reads(np.ascontiguousarray(words, dtype=np.uint32), off, r.length[idx].copy(), r.ids[idx].copy(), r.name_key[idx].copy(), r.name_hash[idx].copy())


def _ndigits(ids):
    ids = ids.astype(np.uint64)
    nd = np.ones(ids.shape, dtype=np.int64)
    p = np.uint64(10)
    for _ in range(9):
        nd += (ids >= p)
        p = p * np.uint64(10)
    return nd


def name_keys(ids):
    """strcmp-order-preserving u64 key of the decimal string of each id: digit d at string position i
    becomes nibble (d+1) at bits [60-4i, 64-4i); shorter strings are zero padded, so a proper prefix
    sorts first exactly as strcmp does (ref: minimap2/map.c:132 strcmp(qname, s->name))."""
    ids = np.asarray(ids, dtype=np.uint64)
    nd = _ndigits(ids)
    key = np.zeros(ids.shape, dtype=np.uint64)
    for i in range(10):
        valid = nd > i
        e = np.where(valid, nd - 1 - i, 0).astype(np.uint64)
        dig = (ids // (np.uint64(10) ** e)) % np.uint64(10)
        key |= np.where(valid, (dig + np.uint64(1)) << np.uint64(60 - 4 * i), np.uint64(0))
    return key


def x31_hash(ids):
    """__ac_X31_hash_string of the decimal string (ref: minimap2/khash.h:383-388)."""
    ids = np.asarray(ids, dtype=np.uint64)
    nd = _ndigits(ids)
    h = np.zeros(ids.shape, dtype=np.uint64)
    for i in range(10):
        valid = nd > i
        e = np.where(valid, nd - 1 - i, 0).astype(np.uint64)
        ch = (ids // (np.uint64(10) ** e)) % np.uint64(10) + np.uint64(48)
        nh = ch if i == 0 else ((h * np.uint64(31) + ch) & np.uint64(0xffffffff))
        h = np.where(valid, nh, h)
    return h.astype(np.uint32)


def make_reads(words_list, lengths, ids):
    lengths = np.asarray(lengths, dtype=np.uint32)
    ids = np.asarray(ids, dtype=np.uint32)
    nw = (lengths.astype(np.uint64) + np.uint64(15)) // np.uint64(16)
    off = np.zeros(len(lengths) + 1, dtype=np.uint64)
    np.cumsum(nw, out=off[1:])
    words = np.concatenate(words_list).astype(np.uint32) if len(words_list) else np.zeros(0, np.uint32)
    assert words.shape[0] == int(off[-1])
    return Reads(np.ascontiguousarray(words), off, lengths, ids, name_keys(ids), x31_hash(ids))


_CODE = np.full(256, 4, dtype=np.uint8)
for _c, _v in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"TtUu", 3)):
    for _ch in _c:
        _CODE[_ch] = _v


def reads_from_ascii(seqs, ids):
    """ASCII sequences (bytes) -> Reads with an ambiguity mask (any non-ACGTU base, ref: minimap2/sketch.c:9-26).
    Each read occupies an even number of words so that the mask offset is word_off/2 (include/ndgpu.h)."""
    words, masks, lens = [], [], []
    any_n = False
    for s in seqs:
        c = _CODE[np.frombuffer(s, dtype=np.uint8)]
        n = c.shape[0]
        amb = c == 4
        any_n |= bool(amb.any())
        w = pack_bases(np.where(amb, 0, c).astype(np.uint8)) if n else np.zeros(0, np.uint32)
        if w.shape[0] & 1:
            w = np.concatenate([w, np.zeros(1, np.uint32)])
        m = np.zeros(w.shape[0] // 2, dtype=np.uint32)
        if n:
            bits = np.zeros(m.shape[0] * 32, dtype=np.uint32); bits[:n] = amb
            m = np.bitwise_or.reduce(bits.reshape(-1, 32) << np.arange(32, dtype=np.uint32), axis=1).astype(np.uint32)
        words.append(w); masks.append(m); lens.append(n)
    ids = np.asarray(ids, dtype=np.uint32)
    lens = np.asarray(lens, dtype=np.uint32)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([w.shape[0] for w in words], out=off[1:])
    allw = np.concatenate(words).astype(np.uint32) if words else np.zeros(0, np.uint32)
    allm = np.concatenate(masks + [np.zeros(2, np.uint32)]).astype(np.uint32)
    return Reads(np.ascontiguousarray(allw), off, lens, ids, name_keys(ids), x31_hash(ids), allm)


def read_fastx(path, max_reads=None):
    """FASTA/FASTQ, optionally gzipped (ref: lib/mseq.h kseq, minimap2/bseq.c:39-49).  Names must be the decimal
    integers seq_dump / the .ovl format work with (ref: minimap2/map.c:1300 strtoul)."""
    import gzip
    with open(path, "rb") as fh:
        magic = fh.read(2)
    op = gzip.open if magic == b"\x1f\x8b" else open
    seqs, ids = [], []
    with op(path, "rb") as fh:
        name, chunks, mode = None, [], None
        qual_left = 0
        for line in fh:
            line = line.rstrip(b"\r\n")
            if qual_left > 0:
                qual_left -= len(line)
                continue
            if line[:1] in (b">", b"@") and (mode != "fq_seq" or line[:1] == b"@") and not (mode == "fq_seq" and chunks == [] and False):
                if name is not None:
                    seqs.append(b"".join(chunks)); ids.append(name)
                    if max_reads is not None and len(seqs) >= max_reads:
                        name = None
                        break
                tok = line[1:].split()[0] if len(line) > 1 else b""
                if not tok.isdigit():
                    raise ValueError("read name %r is not a decimal integer (the .ovl format stores strtoul(name))" % tok)
                name, chunks, mode = int(tok), [], ("fq_seq" if line[:1] == b"@" else "fa")
            elif line[:1] == b"+" and mode == "fq_seq":
                qual_left = sum(len(c) for c in chunks)
            else:
                chunks.append(line)
        if name is not None:
            seqs.append(b"".join(chunks)); ids.append(name)
    return reads_from_ascii(seqs, ids)


def read_seqfile(path, max_reads=None):
    """format sniffing as the reference does (ref: lib/bseq.c:99-112: first byte 0 or 254 = .2bit)"""
    with open(path, "rb") as fh:
        b = fh.read(2)
    if len(b) == 2 and b[0] in (0, 254):
        return read_2bit(path, max_reads)
    return read_fastx(path, max_reads)


def read_2bit(path, max_reads=None):
    """Parse a .2bit container into the C-ABI layout (words copied verbatim, record headers stripped)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.shape[0] < 2 or raw[0] not in (0, 254):
        raise ValueError("%s: not a .2bit file" % path)
    pos, n = 2, raw.shape[0]
    words, lens, ids = [], [], []
    while pos + 8 <= n:
        name, ln = np.frombuffer(raw, dtype="<u4", count=2, offset=pos)
        nw = (int(ln) + 15) // 16
        pos += 8
        if pos + 4 * nw > n:
            break
        words.append(np.frombuffer(raw, dtype="<u4", count=nw, offset=pos))
        lens.append(int(ln)); ids.append(int(name))
        pos += 4 * nw
        if max_reads is not None and len(lens) >= max_reads:
            break
    return make_reads(words, lens, ids)


def write_2bit(path, reads: Reads, idx_path=None):
    """ref: lib/bseq.c:114-139 (seq2bit) + util/seq_dump.c:36-41 (.idx sidecar line)."""
    with open(path, "wb") as fo:
        fo.write(bytes([0, 254]))
        off = 2
        lines = []
        for i in range(reads.n):
            w0, w1 = int(reads.word_off[i]), int(reads.word_off[i + 1])
            fo.write(np.array([reads.ids[i], reads.length[i]], dtype="<u4").tobytes())
            fo.write(reads.words[w0:w1].astype("<u4").tobytes())
            lines.append("%u\t%u\t%u\n" % (reads.ids[i], off + 8, reads.length[i]))
            off += 8 + 4 * (w1 - w0)
    if idx_path:
        with open(idx_path, "w") as fi:
            fi.writelines(lines)


def pack_bases(codes):
    """codes: uint8 array of 0..3 -> uint32 words (first base in the top bits)."""
    n = codes.shape[0]
    nw = (n + 15) // 16
    pad = np.zeros(nw * 16, dtype=np.uint32)
    pad[:n] = codes
    pad = pad.reshape(nw, 16)
    sh = (30 - 2 * np.arange(16, dtype=np.uint32)).astype(np.uint32)
    return np.bitwise_or.reduce(pad << sh, axis=1).astype(np.uint32)


def unpack_bases(words, length):
    sh = (30 - 2 * np.arange(16, dtype=np.uint32)).astype(np.uint32)
    c = ((words[:, None] >> sh[None, :]) & np.uint32(3)).astype(np.uint8).reshape(-1)
    return c[:length]


def decode_ovl8(buf):
    """Decode an 8-field .ovl stream to an (n, 8) uint32 array: qname rev qs qe tname ts te match
    (ref: lib/ovl.c:150-203 decode_ovl; util/ovl_cvt.c:38-40 column order)."""
    b = np.frombuffer(buf, dtype=np.uint8)
    ends = np.nonzero(b < 128)[0]
    starts = np.concatenate(([0], ends[:-1] + 1))
    nv = ends.shape[0]
    if nv % 8:
        raise ValueError("truncated .ovl stream")
    vals = np.zeros(nv, dtype=np.uint64)
    ln = ends - starts + 1
    for k in range(int(ln.max()) if nv else 0):
        sel = ln > k
        vals[sel] = (vals[sel] << np.uint64(7)) | (b[starts[sel] + k] & 127).astype(np.uint64)
    v = vals.reshape(-1, 8).astype(np.int64)
    flags = v[:, 1]
    out = np.zeros((v.shape[0], 8), dtype=np.int64)
    dq = np.where(flags & 2, -v[:, 0], v[:, 0])
    dt = np.where(flags & 4, -v[:, 4], v[:, 4])
    out[:, 0] = np.cumsum(dq)
    out[:, 4] = np.cumsum(dt)
    out[:, 1] = flags & 1
    out[:, 2] = v[:, 2]
    out[:, 3] = v[:, 2] + v[:, 3]
    out[:, 5] = v[:, 5]
    out[:, 6] = np.where(flags & 8, v[:, 5] + v[:, 3] + v[:, 6], v[:, 5] + v[:, 3] - v[:, 6])
    out[:, 7] = v[:, 7]
    return out.astype(np.uint32)


def decode_ovl10(buf):
    """Decode a 10-field .ovl stream (`--step 2`; starts with the bytes 0, 255) to an (n, 10) uint32 array in ovl_cvt column order:
    qname rev qs qe tname ts te qlen tlen identity (ref: lib/ovl.c:150-203 decode_ovl with n = 10, util/ovl_cvt.c:33-50: a zero
    length means "same read as the previous record")."""
    b = np.frombuffer(buf, dtype=np.uint8)
    if b.shape[0] < 2 or b[0] != 0 or b[1] != 255:
        raise ValueError("not a 10-field .ovl stream (header bytes 0, 255 missing)")
    b = b[2:]
    ends = np.nonzero(b < 128)[0]
    starts = np.concatenate(([0], ends[:-1] + 1)) if ends.shape[0] else ends
    nv = ends.shape[0]
    if nv % 10:
        raise ValueError("truncated .ovl stream")
    vals = np.zeros(nv, dtype=np.uint64)
    ln = ends - starts + 1
    for k in range(int(ln.max()) if nv else 0):
        sel = ln > k
        vals[sel] = (vals[sel] << np.uint64(7)) | (b[starts[sel] + k] & 127).astype(np.uint64)
    v = vals.reshape(-1, 10).astype(np.int64)
    flags = v[:, 1]
    out = np.zeros((v.shape[0], 10), dtype=np.int64)
    out[:, 0] = np.cumsum(np.where(flags & 2, -v[:, 0], v[:, 0]))
    out[:, 4] = np.cumsum(np.where(flags & 4, -v[:, 4], v[:, 4]))
    out[:, 1] = flags & 1
    out[:, 2] = v[:, 2]
    out[:, 3] = v[:, 2] + v[:, 3]
    out[:, 5] = v[:, 5]
    out[:, 6] = np.where(flags & 8, v[:, 5] + v[:, 3] + v[:, 6], v[:, 5] + v[:, 3] - v[:, 6])
    for col in (7, 8):                                  # lengths: carried forward while zero
        x = v[:, col]
        idx = np.where(x != 0, np.arange(x.shape[0]), 0)
        np.maximum.accumulate(idx, out=idx)
        out[:, col] = x[idx] if x.shape[0] else x
    out[:, 9] = v[:, 9]
    return out.astype(np.uint32)


def encode_ovl8_py(records):
    """Pure-Python reference encoder for small cases (tests).  records: iterable of 8-tuples in
    ovl_cvt column order.  ref: lib/ovl.c:109-150."""
    out = bytearray()
    pq = pt = 0

    def put(v):
        if v <= 127:
            out.append(v)
            return
        groups = []
        while v:
            groups.append(v & 127)
            v >>= 7
        groups.reverse()
        for g in groups[:-1]:
            out.append(g | 128)
        out.append(groups[-1])

    for (qn, rev, qs, qe, tn, ts, te, match) in records:
        flags = rev
        f0 = qn - pq if qn >= pq else pq - qn
        if qn < pq: flags |= 2
        f4 = tn - pt if tn >= pt else pt - tn
        if tn < pt: flags |= 4
        pq, pt = qn, tn
        ql, tl = qe - qs, te - ts
        f6 = ql - tl if ql >= tl else tl - ql
        if ql < tl: flags |= 8
        for v in (f0, flags, qs, ql, f4, ts, f6, match):
            put(int(v))
    return bytes(out)
