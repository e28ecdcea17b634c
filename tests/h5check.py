"""Test infrastructure for tests/test_hdf5.py -- two independent pure-Python restatements of the HDF5 file format
specification (version 3.0), written apart from csrc/hdf5io.cpp so that reader and writer are not only checked against
each other:

* ``parse_v0``  walks a file the way libhdf5 does for the "earliest" format (superblock 0, symbol-table groups, version-1
  object headers) and asserts the structural invariants libhdf5 relies on (B-tree key order, sibling links, heap offsets,
  node capacities, end-of-file address).  Used on the WRITER's output.
* ``build_latest``  emits a file in the encodings libhdf5 uses with ``libver='latest'``-style objects (superblock 2,
  version-2 object headers with continuation, link messages, dataspace 2, compound 3, attribute 3, layout 3 compact /
  chunked with a version-1 chunk B-tree, deflate + shuffle, big-endian and f32 members).  Used as input of the READER.
"""
import struct
import zlib

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


# ------------------------------------------------------------------ parser (earliest format)
def _u(b, off, n):
    return int.from_bytes(b[off:off + n], "little")


def _messages_v1(b, addr):
    assert b[addr] == 1, "object header version"
    nmsgs = _u(b, addr + 2, 2)
    size = _u(b, addr + 8, 4)
    p, end, out = addr + 16, addr + 16 + size, []
    while p < end:
        t, s, fl = _u(b, p, 2), _u(b, p + 2, 2), b[p + 4]
        assert s % 8 == 0, "message size must be a multiple of 8"
        out.append((t, fl, b[p + 8:p + 8 + s]))
        p += 8 + s
    assert p == end and len(out) == nmsgs, "header size / message count disagree"
    return out


def _heap(b, addr):
    assert b[addr:addr + 4] == b"HEAP" and b[addr + 4] == 0
    size, free, seg = _u(b, addr + 8, 8), _u(b, addr + 16, 8), _u(b, addr + 24, 8)
    assert free == 1 or free + 16 <= size
    assert seg + size <= len(b)
    assert b[seg:seg + 8] == b"\0" * 8, "offset 0 of a group heap is the empty string"
    return b[seg:seg + size]


def _name(heap, off):
    assert off % 8 == 0 and off < len(heap)
    return heap[off:heap.index(b"\0", off)].decode()


def _walk_tree(b, addr, heap, leaf_k, internal_k, expect_left=UNDEF):
    """returns (list of (name, header, cache, scratch)), level, first_key, last_key"""
    assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 0
    level, used = b[addr + 5], _u(b, addr + 6, 2)
    assert used <= 2 * internal_k
    left, right = _u(b, addr + 8, 8), _u(b, addr + 16, 8)
    p = addr + 24
    keys, kids = [], []
    for i in range(used):
        keys.append(_u(b, p, 8)); kids.append(_u(b, p + 8, 8)); p += 16
    keys.append(_u(b, p, 8))
    names = [_name(heap, k) for k in keys]
    assert names == sorted(names), "B-tree keys must ascend"
    entries = []
    for i, kid in enumerate(kids):
        if level == 0:
            assert b[kid:kid + 4] == b"SNOD" and b[kid + 4] == 1
            n = _u(b, kid + 6, 2)
            assert 1 <= n <= 2 * leaf_k
            sub = []
            for q in range(n):
                e = kid + 8 + 40 * q
                sub.append((_name(heap, _u(b, e, 8)), _u(b, e + 8, 8), _u(b, e + 16, 4), b[e + 24:e + 40]))
        else:
            sub, lvl, _, _ = _walk_tree(b, kid, heap, leaf_k, internal_k)
            assert lvl == level - 1
        ns = [s[0] for s in sub]
        assert ns == sorted(ns) and len(set(ns)) == len(ns)
        # child i holds the names in (key[i], key[i+1]]
        assert names[i] < ns[0] or (names[i] == "" and i == 0 and ns[0] >= ""), (names[i], ns[0])
        assert ns[-1] == names[i + 1], "right key of a child is its largest name"
        entries += sub
    return entries, level, left, right


def _siblings_ok(b, root, internal_k):
    """every level is a doubly linked list from left to right"""
    level_nodes = [root]
    while True:
        for i, a in enumerate(level_nodes):
            left, right = _u(b, a + 8, 8), _u(b, a + 16, 8)
            assert left == (level_nodes[i - 1] if i else UNDEF)
            assert right == (level_nodes[i + 1] if i + 1 < len(level_nodes) else UNDEF)
        if b[level_nodes[0] + 5] == 0:
            return
        nxt = []
        for a in level_nodes:
            used = _u(b, a + 6, 2)
            nxt += [_u(b, a + 24 + 16 * i + 8, 8) for i in range(used)]
        level_nodes = nxt


def _group(b, header, leaf_k, internal_k):
    msgs = _messages_v1(b, header)
    st = [m for m in msgs if m[0] == 0x11]
    assert len(st) == 1
    btree, heap_addr = _u(st[0][2], 0, 8), _u(st[0][2], 8, 8)
    heap = _heap(b, heap_addr)
    entries, _, left, right = _walk_tree(b, btree, heap, leaf_k, internal_k)
    assert left == UNDEF and right == UNDEF
    _siblings_ok(b, btree, internal_k)
    return entries, btree, heap_addr


def _datatype(d):
    cls, ver = d[0] & 15, d[0] >> 4
    size = _u(d, 4, 4)
    if cls == 1:
        assert ver == 1 and d[1:4] == b"\x20\x3f\x00" and size == 8
        assert d[8:20] == struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return ("f64",), 20
    if cls == 0:
        assert ver == 1 and size == 8 and _u(d, 8, 2) == 0 and _u(d, 10, 2) == 64
        return ("i64" if d[1] & 8 else "u64",), 12
    assert cls == 6 and ver == 1
    n = _u(d, 1, 2)
    p, members = 8, []
    for _ in range(n):
        end = d.index(b"\0", p)
        name = d[p:end].decode()
        p += (end - p + 1 + 7) // 8 * 8
        off = _u(d, p, 4)
        assert d[p + 4] == 0 and d[p + 5:p + 32] == b"\0" * 27, "scalar member: dimensionality 0, rest reserved"
        p += 32
        t, used = _datatype(d[p:])
        p += used
        members.append((name, off, t))
    return ("compound", size, members), p


def parse_v0(b):
    """-> {name: dict(shape, bids, data_addr, data_bytes)} for /tensors; asserts the invariants on the way."""
    assert b[:8] == SIG
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k = _u(b, 16, 2), _u(b, 18, 2)
    assert _u(b, 24, 8) == 0 and _u(b, 32, 8) == UNDEF and _u(b, 48, 8) == UNDEF
    assert _u(b, 40, 8) == len(b), "end-of-file address"
    assert _u(b, 56, 8) == 0
    root_header, cache = _u(b, 64, 8), _u(b, 72, 4)
    root, r_btree, r_heap = _group(b, root_header, leaf_k, internal_k)
    assert cache == 1 and (_u(b, 80, 8), _u(b, 88, 8)) == (r_btree, r_heap), "root entry caches its B-tree and heap"
    assert [e[0] for e in root] == ["tensors"]
    _, t_header, t_cache, t_scratch = root[0]
    members, t_btree, t_heap = _group(b, t_header, leaf_k, internal_k)
    assert t_cache == 1 and (_u(t_scratch, 0, 8), _u(t_scratch, 8, 8)) == (t_btree, t_heap)
    out = {}
    for name, header, c, _ in members:
        assert c == 0
        info = {"bids": None}
        seen = set()
        for t, fl, d in _messages_v1(b, header):
            seen.add(t)
            if t == 0x01:
                assert d[0] == 1 and d[2] == 0
                info["shape"] = [_u(d, 8 + 8 * i, 8) for i in range(d[1])]
            elif t == 0x03:
                ty, _ = _datatype(d)
                assert ty == ("compound", 16, [("re", 0, ("f64",)), ("im", 8, ("f64",))]), ty
            elif t == 0x05:
                assert d[:8] == bytes([2, 2, 2, 1, 0, 0, 0, 0])
            elif t == 0x08:
                assert d[0] == 3 and d[1] == 1
                info["data_addr"], info["data_bytes"] = _u(d, 2, 8), _u(d, 10, 8)
            elif t == 0x0C:
                assert d[0] == 1
                nsz, tsz, ssz = _u(d, 2, 2), _u(d, 4, 2), _u(d, 6, 2)
                p = 8
                nm = d[p:p + nsz]; p += (nsz + 7) // 8 * 8
                ty, used = _datatype(d[p:p + tsz]); assert used == tsz; p += (tsz + 7) // 8 * 8
                sp = d[p:p + ssz]; p += (ssz + 7) // 8 * 8
                assert nm == b"bids\0" and ty == ("u64",) and sp[0] == 1 and sp[1] == 1
                cnt = _u(sp, 8, 8)
                info["bids"] = [_u(d, p + 8 * i, 8) for i in range(cnt)]
        assert {1, 3, 8} <= seen
        n = 1
        for s in info["shape"]:
            n *= s
        assert info["data_bytes"] == 16 * n
        if info["data_addr"] != UNDEF:
            assert info["data_addr"] + info["data_bytes"] <= len(b)
        out[name] = info
    return out


# ------------------------------------------------------------------ builder ("latest"-style encodings)
def _f_type(size=8, big=False):
    bits0 = 0x20 | (1 if big else 0)
    if size == 8:
        return bytes([0x11, bits0, 63, 0]) + struct.pack("<IHHBBBBI", 8, 0, 64, 52, 11, 0, 52, 1023)
    return bytes([0x11, bits0, 31, 0]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)


def _complex_type_v3(names=("r", "i"), fsize=8, big=False):
    size = 2 * fsize
    out = bytes([0x36, 2, 0, 0]) + struct.pack("<I", size)
    for i, n in enumerate(names):
        out += n.encode() + b"\0" + bytes([i * fsize]) + _f_type(fsize, big)     # offset: 1 byte while size < 256
    return out


def _int_type(size, signed, big=False):
    return bytes([0x10, (8 if signed else 0) | (1 if big else 0), 0, 0]) + struct.pack("<IHH", size, 0, 8 * size)


def _space_v2(shape):
    if shape is None:
        return bytes([2, 0, 0, 2])                     # null
    return bytes([2, len(shape), 0, 1 if shape else 0]) + b"".join(struct.pack("<Q", s) for s in shape)


def _msg_v2(t, body, flags=0):
    return bytes([t]) + struct.pack("<H", len(body)) + bytes([flags]) + body


def _attr_v3(name, values, size, signed, big=False):
    ty, sp = _int_type(size, signed, big), _space_v2([len(values)])
    data = b"".join(int(v).to_bytes(size, "big" if big else "little", signed=signed) for v in values)
    nm = name.encode() + b"\0"
    return bytes([3, 0]) + struct.pack("<HHH", len(nm), len(ty), len(sp)) + b"\0" + nm + ty + sp + data


class LatestFile:
    """Collects blocks at increasing addresses; objects are version-2 headers."""

    def __init__(self, userblock=0):
        self.base = userblock
        self.buf = bytearray(48)                       # superblock 2: 8 + 4 + 4 * 8 + 4

    def alloc(self, data):
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, msgs, split_at=None):
        """msgs: list of encoded version-2 messages.  split_at: put messages[split_at:] into a continuation block."""
        if split_at is None:
            body = b"".join(msgs)
            return self.alloc(b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body)) + body + b"\0\0\0\0")
        tail = b"".join(msgs[split_at:])
        cont = self.alloc(b"OCHK" + tail + b"\0\0\0\0")
        body = b"".join(msgs[:split_at]) + _msg_v2(0x10, struct.pack("<QQ", cont, len(tail) + 8))
        return self.alloc(b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body)) + body + b"\0\0\0\0")

    def group(self, links):
        msgs = [_msg_v2(0x02, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF)),      # link info: compact storage
                _msg_v2(0x0A, bytes([0, 0]))]                                          # group info
        for name, addr in links:
            nm = name.encode()
            msgs.append(_msg_v2(0x06, bytes([1, 0x00, len(nm)]) + nm + struct.pack("<Q", addr)))
        return self.object_header(msgs)

    def dataset(self, shape, dtype, layout_msg, attrs=(), filters=None, split=False):
        msgs = [_msg_v2(0x01, _space_v2(shape)), _msg_v2(0x03, dtype, 1), _msg_v2(0x05, bytes([3, 0x09])), _msg_v2(0x08, layout_msg)]
        if filters is not None:
            msgs.append(_msg_v2(0x0B, filters))
        msgs += [_msg_v2(0x0C, a) for a in attrs]
        return self.object_header(msgs, split_at=3 if split else None)

    def contiguous(self, data):
        addr = self.alloc(data)
        return bytes([3, 1]) + struct.pack("<QQ", addr, len(data))

    @staticmethod
    def compact(data):
        return bytes([3, 0]) + struct.pack("<H", len(data)) + data

    def chunked(self, shape, chunk, esize, chunks):
        """chunks: list of (offsets, raw bytes, filter mask); one leaf B-tree node."""
        addrs = [self.alloc(raw) for _, raw, _ in chunks]
        node = b"TREE" + bytes([1, 0]) + struct.pack("<HQQ", len(chunks), UNDEF, UNDEF)
        for (off, raw, mask), a in zip(chunks, addrs):
            node += struct.pack("<II", len(raw), mask) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0)
            node += struct.pack("<Q", a)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in shape) + struct.pack("<Q", 0)
        bt = self.alloc(node)
        return bytes([3, 2, len(shape) + 1]) + struct.pack("<Q", bt) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", esize)

    def finish(self, root_group):
        sb = SIG + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", self.base, UNDEF, len(self.buf), root_group) + b"\0\0\0\0"
        self.buf[:48] = sb
        return b"\0" * self.base + bytes(self.buf)


def filters_v2(deflate=True, shuffle_size=None):
    fl = []
    if shuffle_size:
        fl.append(struct.pack("<HHHI", 2, 0, 1, shuffle_size))
    if deflate:
        fl.append(struct.pack("<HHHI", 1, 0, 1, 6))
    return bytes([2, len(fl)]) + b"".join(fl)


def shuffle(raw, esize):
    n = len(raw) // esize
    return bytes(raw[i * esize + b] for b in range(esize) for i in range(n))


def deflate(raw):
    return zlib.compress(raw, 6)
