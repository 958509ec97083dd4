"""TEST INFRASTRUCTURE — an independent Python restatement of the reference's `.impg` index file
(Impg::serialize_with_forest_map / load_from_file, reference src/impg.rs:1655-1850) over serde +
bincode 2 `config::standard()` (little endian, variable-length integers, zigzag for signed,
length-prefixed strings / sequences / maps, struct fields in declaration order).

PARITY UNPINNED: bincode 2.0.1 is not vendored (Cargo.lock:186-189) and the reference holds no .impg
fixture; this module and impg_b200/csrc/impg_file.cu restate the published encoding separately and the
tests require them to agree byte for byte."""
import struct

STRAND_BIT = 1 << 63    # src/impg.rs:177
REVERSED_BIT = 1 << 62  # src/impg.rs:178
OFFSET_MASK = ~(STRAND_BIT | REVERSED_BIT) & (2**64 - 1)


def varint(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + struct.pack("<H", v)
    if v < 1 << 32:
        return b"\xfc" + struct.pack("<I", v)
    if v < 1 << 64:
        return b"\xfd" + struct.pack("<Q", v)
    raise ValueError(v)


def zigzag(v):
    return varint((v << 1) if v >= 0 else ((-v) << 1) - 1)


def string(s):
    b = s.encode()
    return varint(len(b)) + b


def parse_paf_like_reference(paths):
    """src/paf.rs:118-194 per file, shared SequenceIndex (ids by first appearance, query column first):
    returns names, lens, records [(q, t, qs, qe, ts, te, strand, file, cg_offset, cg_bytes)]."""
    names, lens, ids, recs = [], [], {}, []

    def get(name, length):
        if name not in ids:
            ids[name] = len(names)
            names.append(name)
            lens.append(length)
        return ids[name]

    for fi, path in enumerate(paths):
        pos = 0
        for raw in open(path, "rb").read().split(b"\n"):
            line = raw[:-1] if raw.endswith(b"\r") else raw
            if not line and not raw:
                continue
            f = line.split(b"\t")
            q = get(f[0].decode(), int(f[1]))
            t = get(f[5].decode(), int(f[6]))
            off, nbytes = pos, 0
            for tag in f:
                if tag.startswith(b"cg:Z:"):
                    off += 5
                    nbytes = len(tag) - 5
                    break
                off += len(tag) + 1
            recs.append((q, t, int(f[2]), int(f[3]), int(f[7]), int(f[8]), 1 if f[4][:1] == b"-" else 0, fi, off, nbytes))
            pos += len(line) + 1
    return names, lens, recs


def entries_by_target(recs, bidirectional=True):
    """from_multi_alignment_records (src/impg.rs:1562-1605): forward entry under the target, reversed copy
    under the query (not for self alignments); a tree's in-order dump = stable sort by `first`."""
    trees = {}
    for q, t, qs, qe, ts, te, strand, fi, off, nbytes in recs:
        sdo = off | (STRAND_BIT if strand else 0)
        trees.setdefault(t, []).append((ts, te, q, ts, te, qs, qe, fi, sdo, nbytes))
        if bidirectional and q != t:
            trees.setdefault(q, []).append((qs, qe, t, qs, qe, ts, te, fi, sdo | REVERSED_BIT, nbytes))
    return {t: sorted(v, key=lambda e: e[0]) for t, v in trees.items()}


def encode(names, lens, trees, map_order=None, tree_order=None, magic=b"IMPGIDX2"):
    """The file bytes. map_order / tree_order: iteration orders of the hash maps (any order is a valid file)."""
    ids = list(range(len(names))) if map_order is None else list(map_order)
    out = bytearray(magic + b"\0" * 8)
    out += varint(len(ids)) + b"".join(string(names[i]) + varint(i) for i in ids)
    out += varint(len(ids)) + b"".join(varint(i) + string(names[i]) for i in ids)
    out += varint(len(ids)) + b"".join(varint(i) + varint(lens[i]) for i in ids)
    out += varint(len(names))
    forest = []
    for t in (sorted(trees) if tree_order is None else tree_order):
        forest.append((t, len(out)))
        out += varint(t) + varint(len(trees[t]))
        for first, last, q, ts, te, qs, qe, fi, sdo, nbytes in trees[t]:
            out += zigzag(first) + zigzag(last) + varint(q) + zigzag(ts) + zigzag(te) + zigzag(qs) + zigzag(qe)
            out += varint(fi) + varint(sdo) + varint(nbytes)
    fo = len(out)
    out += varint(len(forest)) + b"".join(varint(t) + varint(o) for t, o in forest)
    out[8:16] = struct.pack("<Q", fo)
    return bytes(out)


class _R:
    def __init__(self, b, p):
        self.b, self.p = b, p

    def varint(self):
        c = self.b[self.p]
        self.p += 1
        if c < 251:
            return c
        n = {251: 2, 252: 4, 253: 8}[c]
        v = int.from_bytes(self.b[self.p:self.p + n], "little")
        self.p += n
        return v

    def zigzag(self):
        z = self.varint()
        return (z >> 1) ^ -(z & 1)

    def string(self):
        n = self.varint()
        s = self.b[self.p:self.p + n].decode()
        self.p += n
        return s


def decode(data):
    """-> (names by id, lens by id, {target: [entry tuples]})."""
    assert data[:8] in (b"IMPGIDX2", b"IMPGIDX1")
    fo = struct.unpack("<Q", data[8:16])[0]
    r = _R(data, 16)
    n2i = {}
    for _ in range(r.varint()):
        s = r.string()
        n2i[s] = r.varint()
    i2n = {}
    for _ in range(r.varint()):
        i = r.varint()
        i2n[i] = r.string()
    i2l = {}
    for _ in range(r.varint()):
        i = r.varint()
        i2l[i] = r.varint()
    nxt = r.varint()
    assert {v: k for k, v in n2i.items()} == i2n
    fr = _R(data, fo)
    forest = {}
    for _ in range(fr.varint()):
        t = fr.varint()
        forest[t] = fr.varint()
    trees = {}
    for t, off in forest.items():
        tr = _R(data, off)
        assert tr.varint() == t
        ent = []
        for _ in range(tr.varint()):
            ent.append((tr.zigzag(), tr.zigzag(), tr.varint(), tr.zigzag(), tr.zigzag(), tr.zigzag(), tr.zigzag(),
                        tr.varint(), tr.varint(), tr.varint()))
        trees[t] = ent
    return [i2n.get(i, "") for i in range(nxt)], [i2l.get(i, 0) for i in range(nxt)], trees


# ---------------------------------------------------------------- BGZF (src/paf.rs:47-114, :199-302)
def bgzf_compress(data, block=700):
    """`data` as a BGZF file with small blocks (so that lines and CIGARs straddle block borders) plus the EOF
    marker block; returns (file bytes, [(compressed offset, inflated start, inflated length)])."""
    import zlib

    out, table, u = bytearray(), [], 0
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
    for ch in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(ch) + co.flush()
        bsize = 18 + len(body) + 8
        hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        table.append((len(out), u, len(ch)))
        out += hdr + body + struct.pack("<II", zlib.crc32(ch) & 0xffffffff, len(ch))
        u += len(ch)
    return bytes(out), table


def virtual_position(table, uoff):
    """Virtual position of inflated offset `uoff`: the block that holds the byte (noodles reports the end of a block as
    offset 0 of the next one)."""
    for coff, ustart, ulen in table:
        if ustart <= uoff < ustart + ulen:
            return (coff << 16) | (uoff - ustart)
    raise ValueError(uoff)
