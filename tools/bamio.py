"""Tiny pure-Python BGZF/BAM reader + writer for tests and fixture preparation (no pysam/htslib here).

Only what the tests need: iterate records, look at / rewrite aux tags, write a coordinate-sorted BAM
back out (BGZF members of <= 64 KiB, EOF marker). Not used by the product path.
"""
import struct
import zlib

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def read_bgzf(path):
    data = open(path, "rb").read()
    out = []
    off = 0
    while off < len(data):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        x = off + 12
        bsize = None
        while x < off + 12 + xlen:
            si1, si2, slen = data[x], data[x + 1], struct.unpack_from("<H", data, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", data, x + 4)[0]
            x += 4 + slen
        clen = bsize + 1
        out.append(zlib.decompress(data[off + 12 + xlen: off + clen - 8], -15))
        off += clen
    return b"".join(out)


class Bam:
    def __init__(self, path=None):
        self.header_text = b""
        self.refs = []      # (name, length)
        self.records = []   # raw record bytes (without block_size)
        if path:
            self.load(path)

    def load(self, path):
        raw = read_bgzf(path)
        assert raw[:4] == b"BAM\x01"
        l_text = struct.unpack_from("<I", raw, 4)[0]
        self.header_text = raw[8:8 + l_text]
        o = 8 + l_text
        n_ref = struct.unpack_from("<I", raw, o)[0]
        o += 4
        for _ in range(n_ref):
            ln = struct.unpack_from("<I", raw, o)[0]
            name = raw[o + 4:o + 4 + ln - 1].decode()
            length = struct.unpack_from("<I", raw, o + 4 + ln)[0]
            self.refs.append((name, length))
            o += 8 + ln
        while o + 4 <= len(raw):
            bs = struct.unpack_from("<I", raw, o)[0]
            self.records.append(raw[o + 4:o + 4 + bs])
            o += 4 + bs

    def write(self, path, level=6):
        parts = [b"BAM\x01", struct.pack("<I", len(self.header_text)), self.header_text, struct.pack("<I", len(self.refs))]
        for name, length in self.refs:
            nb = name.encode() + b"\x00"
            parts += [struct.pack("<I", len(nb)), nb, struct.pack("<I", length)]
        for r in self.records:
            parts += [struct.pack("<I", len(r)), r]
        raw = b"".join(parts)
        with open(path, "wb") as f:
            for i in range(0, len(raw), 0xff00):
                chunk = raw[i:i + 0xff00]
                co = zlib.compressobj(level, zlib.DEFLATED, -15)
                comp = co.compress(chunk) + co.flush()
                bsize = len(comp) + 25
                f.write(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
                f.write(comp)
                f.write(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
            f.write(BGZF_EOF)


def rec_fields(r):
    tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", r, 0)
    o = 32
    name = r[o:o + l_name - 1].decode()
    o += l_name
    cigar = list(struct.unpack_from("<%dI" % n_cig, r, o))
    o += 4 * n_cig
    seq = r[o:o + (l_seq + 1) // 2]
    o += (l_seq + 1) // 2
    qual = r[o:o + l_seq]
    o += l_seq
    return dict(tid=tid, pos=pos, flag=flag, l_seq=l_seq, name=name, cigar=cigar, seq=seq, qual=qual, aux_off=o)


def iter_aux(r, o):
    """yields (tag, type, start, end) with start/end spanning the whole field"""
    n = len(r)
    while o + 3 <= n:
        tag = r[o:o + 2]
        ty = chr(r[o + 2])
        p = o + 3
        if ty in "AcC":
            e = p + 1
        elif ty in "sS":
            e = p + 2
        elif ty in "iIf":
            e = p + 4
        elif ty in "ZH":
            e = r.index(b"\x00", p) + 1
        elif ty == "B":
            sub = chr(r[p])
            cnt = struct.unpack_from("<I", r, p + 1)[0]
            es = 1 if sub in "cC" else 2 if sub in "sS" else 4
            e = p + 5 + es * cnt
        else:
            raise ValueError("bad aux type " + ty)
        yield tag, ty, o, e
        o = e


def get_aux(r, tag):
    f = rec_fields(r)
    for t, ty, s, e in iter_aux(r, f["aux_off"]):
        if t == tag:
            if ty == "Z":
                return r[s + 3:e - 1]
            if ty == "B":
                return r[s + 8:e]
            return r[s + 3:e]
    return None


def replace_aux(r, mapping):
    """mapping: old tag (bytes) -> (new tag, type char, payload bytes) or None to drop"""
    f = rec_fields(r)
    out = [r[:f["aux_off"]]]
    for t, ty, s, e in iter_aux(r, f["aux_off"]):
        if t in mapping:
            new = mapping[t]
            if new is None:
                continue
            if callable(new):
                new = new(ty, r[s + 3:e])
            ntag, nty, payload = new
            out.append(ntag + nty.encode() + payload)
        else:
            out.append(r[s:e])
    return b"".join(out)
