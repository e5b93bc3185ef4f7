"""`fst` 0.3.x on-disk format, version 2: reader (`get`, `stream`) and builder.

ORACLE / TEST INFRASTRUCTURE ONLY.  The algorithm lives in the third-party crate
`fst` (requirement "0.3.5" at /root/reference/Cargo.toml:23; no Cargo.lock, crate source is
not vendored and not on this box).  This file restates its published format and builder
(SURVEY.md Appendix B).  The reference's call sites are `fst::Map::from_bytes/from_path`
(src/commands/prot2kmer2lca.rs:109-114), `Map::get` (:176; prot2tryp2lca.rs:130),
`Map::stream` (printindex.rs:44-47) and `MapBuilder::{new,insert,finish}`
(buildindex.rs:38-45).

PARITY UNPINNED for byte-compatibility with files written by the real crate: the only
vector the reference holds at this boundary is the `AAAAA->2759, BBBBBB->9153` round trip
(buildindex.rs:20-28).  Everything here is pinned by own writer<->reader round trips only.

Layout summary.  File = u64le version(2), u64le type(0 = Map), nodes..., u64le len,
u64le root_addr.  A node's address is the index of its LAST byte (the state byte); fields
lie before it and are decoded backwards.  Address 0 = the implicit empty final node.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

VERSION = 2
EMPTY_ADDRESS = 0
NONE_ADDRESS = 1
TRANS_INDEX_THRESHOLD = 32

# Frequency-ranked common input bytes of the crate (first 84; only the first 63 are
# encodable in the 6-bit field of the one-transition states).  [recalled]
COMMON_INPUTS_INV = (
    b"te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHGWUV,YKJZXQ;)(~[]$!'*@"
)
_COMMON_IDX = {b: i for i, b in enumerate(COMMON_INPUTS_INV)}


def common_idx(inp: int) -> int:
    """0 = not encodable; else index+1 (<= 63)."""
    i = _COMMON_IDX.get(inp)
    if i is None:
        return 0
    v = (i + 1) % 256
    return v if v <= 0x3F else 0


def pack_size(n: int) -> int:
    s = 1
    while n >= (1 << (8 * s)):
        s += 1
    return s


def _pack(n: int, size: int) -> bytes:
    return n.to_bytes(size, "little")


def _unpack(data: bytes, at: int, size: int) -> int:
    return int.from_bytes(data[at:at + size], "little")


class FstFormatError(Exception):
    pass


# --------------------------------------------------------------------------- reader


class Node:
    """A decoded node: sorted transitions [(input, output, target)], final flag and output."""

    __slots__ = ("final", "final_output", "trans", "_index")

    def __init__(self, final: bool, final_output: int, trans: List[Tuple[int, int, int]]):
        self.final = final
        self.final_output = final_output
        self.trans = trans
        self._index = {t[0]: t for t in trans}

    def find(self, b: int) -> Optional[Tuple[int, int, int]]:
        return self._index.get(b)


def decode_node(data: bytes, a: int) -> Node:
    if a == EMPTY_ADDRESS:
        return Node(True, 0, [])
    s = data[a]
    kind = s >> 6
    if kind == 0b11:  # OneTransNext
        c = s & 0x3F
        il = 1 if c == 0 else 0
        inp = data[a - 1] if il else COMMON_INPUTS_INV[c - 1]
        return Node(False, 0, [(inp, 0, a - il - 1)])
    if kind == 0b10:  # OneTrans
        c = s & 0x3F
        il = 1 if c == 0 else 0
        inp = data[a - 1] if il else COMMON_INPUTS_INV[c - 1]
        z = data[a - il - 1]
        tsz, osz = z >> 4, z & 15
        dpos = a - il - 1 - tsz
        delta = _unpack(data, dpos, tsz)
        start = dpos - osz
        out = _unpack(data, start, osz) if osz else 0
        target = start - delta if delta else EMPTY_ADDRESS
        return Node(False, 0, [(inp, out, target)])
    # AnyTrans
    final = bool(s & 0x40)
    n = s & 0x3F
    nl = 1 if n == 0 else 0
    if nl:
        n = data[a - 1]
        if n == 1:
            n = 256
    base = a - nl - 1
    z = data[base]
    tsz, osz = z >> 4, z & 15
    isz = 256 if n > TRANS_INDEX_THRESHOLD else 0
    start = base - isz - n - n * tsz - n * osz - (osz if final else 0)
    trans = []
    for i in range(n):
        inp = data[base - isz - 1 - i]
        delta = _unpack(data, base - isz - n - (i + 1) * tsz, tsz)
        out = _unpack(data, base - isz - n - n * tsz - (i + 1) * osz, osz) if osz else 0
        trans.append((inp, out, start - delta if delta else EMPTY_ADDRESS))
    fo = _unpack(data, start, osz) if (final and osz) else 0
    return Node(final, fo, trans)


class Fst:
    def __init__(self, data: bytes):
        if len(data) < 32:
            raise FstFormatError("fst too short")
        version, ty = struct.unpack_from("<QQ", data, 0)
        if version == 0 or version > VERSION:
            raise FstFormatError(f"unsupported fst version {version}")
        self.len, self.root_addr = struct.unpack_from("<QQ", data, len(data) - 16)
        if not ((self.root_addr == EMPTY_ADDRESS and len(data) == 32)
                or self.root_addr + 17 == len(data)):
            raise FstFormatError("root address inconsistent with file length")
        self.data = data
        self.type = ty

    def get(self, key: bytes) -> Optional[int]:
        """fst::Map::get: one node hop per key byte, outputs summed."""
        node = decode_node(self.data, self.root_addr)
        out = 0
        for b in key:
            t = node.find(b)
            if t is None:
                return None
            out += t[1]
            node = decode_node(self.data, t[2])
        if not node.final:
            return None
        return out + node.final_output

    def stream(self) -> Iterator[Tuple[bytes, int]]:
        """fst::Map::stream: in-order DFS, one item per final node."""
        stack = [(self.root_addr, b"", 0)]
        # iterative pre-order with explicit reversed push keeps keys ascending
        while stack:
            addr, key, out = stack.pop()
            node = decode_node(self.data, addr)
            if node.final:
                yield key, out + node.final_output
            for inp, o, tgt in reversed(node.trans):
                stack.append((tgt, key + bytes([inp]), out + o))


# --------------------------------------------------------------------------- builder


class _BNode:
    __slots__ = ("is_final", "final_output", "trans", "last")

    def __init__(self):
        self.is_final = False
        self.final_output = 0
        self.trans: List[List[int]] = []   # [inp, out, addr]
        self.last: Optional[List[int]] = None  # [inp, out]

    def last_compiled(self, addr: int):
        if self.last is not None:
            self.trans.append([self.last[0], self.last[1], addr])
            self.last = None

    def add_output_prefix(self, prefix: int):
        if self.is_final:
            self.final_output += prefix
        for t in self.trans:
            t[1] += prefix
        if self.last is not None:
            self.last[1] += prefix

    def key(self):
        return (self.is_final, self.final_output, tuple(tuple(t) for t in self.trans))


class Builder:
    """fst::raw::Builder restated: Daciuk-style incremental construction with the crate's
    bounded (10 000 x 2, FNV-1a) registry for suffix sharing."""

    TABLE_SIZE = 10_000

    def __init__(self):
        self.buf = bytearray(struct.pack("<QQ", VERSION, 0))
        self.stack = [_BNode()]
        self.last_key: Optional[bytes] = None
        self.last_addr = NONE_ADDRESS
        self.len = 0
        self.registry: Dict[int, list] = {}

    # -- registry (Registry::new(10_000, 2), RegistryCache for 2 cells)
    def _hash(self, node: _BNode) -> int:
        M = (1 << 64) - 1
        P = 1099511628211
        h = 14695981039346656037
        h = ((h ^ int(node.is_final)) * P) & M
        h = ((h ^ node.final_output) * P) & M
        for inp, out, addr in node.trans:
            h = ((h ^ inp) * P) & M
            h = ((h ^ out) * P) & M
            h = ((h ^ addr) * P) & M
        return h % self.TABLE_SIZE

    def _compile(self, node: _BNode) -> int:
        if node.is_final and not node.trans and node.final_output == 0:
            return EMPTY_ADDRESS
        cells = self.registry.setdefault(self._hash(node), [None, None])
        k = node.key()
        if cells[0] is not None and cells[0][0] == k:
            return cells[0][1]
        if cells[1] is not None and cells[1][0] == k:
            cells[0], cells[1] = cells[1], cells[0]
            return cells[0][1]
        start = len(self.buf)
        self._emit(node, start)
        self.last_addr = len(self.buf) - 1
        cells[1] = (k, self.last_addr)
        cells[0], cells[1] = cells[1], cells[0]
        return self.last_addr

    def _emit(self, node: _BNode, addr: int):
        buf = self.buf
        if len(node.trans) != 1 or node.is_final:
            n = len(node.trans)
            tsize = 0
            osize = pack_size(node.final_output)
            any_outs = node.final_output != 0
            for inp, out, taddr in node.trans:
                d = 0 if taddr == EMPTY_ADDRESS else addr - taddr
                tsize = max(tsize, pack_size(d))
                osize = max(osize, pack_size(out))
                any_outs = any_outs or out != 0
            if not any_outs:
                osize = 0
            if any_outs:
                if node.is_final:
                    buf += _pack(node.final_output, osize)
                for inp, out, taddr in reversed(node.trans):
                    buf += _pack(out, osize)
            for inp, out, taddr in reversed(node.trans):
                buf += _pack(0 if taddr == EMPTY_ADDRESS else addr - taddr, tsize)
            for inp, out, taddr in reversed(node.trans):
                buf.append(inp)
            if n > TRANS_INDEX_THRESHOLD:
                index = bytearray([255] * 256)
                for i, t in enumerate(node.trans):
                    index[t[0]] = i
                buf += index
            buf.append((tsize << 4) | osize)
            state = 0x40 if node.is_final else 0
            if 1 <= n <= 0x3F:
                state |= n
            else:
                buf.append(1 if n == 256 else n)
            buf.append(state)
            return
        inp, out, taddr = node.trans[0]
        c = common_idx(inp)
        if taddr == self.last_addr and out == 0:
            if c == 0:
                buf.append(inp)
            buf.append(0xC0 | c)
            return
        osize = 0
        if out != 0:
            osize = pack_size(out)
            buf += _pack(out, osize)
        d = 0 if taddr == EMPTY_ADDRESS else addr - taddr
        tsize = pack_size(d)
        buf += _pack(d, tsize)
        buf.append((tsize << 4) | osize)
        if c == 0:
            buf.append(inp)
        buf.append(0x80 | c)

    def _compile_from(self, istate: int):
        addr = NONE_ADDRESS
        while istate + 1 < len(self.stack):
            node = self.stack.pop()
            if addr != NONE_ADDRESS:
                node.last_compiled(addr)
            addr = self._compile(node)
        self.stack[-1].last_compiled(addr) if addr != NONE_ADDRESS else None

    def insert(self, key: bytes, value: int):
        if self.last_key is not None:
            if key == self.last_key:
                raise ValueError("duplicate key")
            if key < self.last_key:
                raise ValueError("keys out of order")
        self.last_key = bytes(key)
        if len(key) == 0:
            self.len = 1
            self.stack[0].is_final = True
            self.stack[0].final_output = value
            return
        out = value
        i = 0
        while i < len(key):
            node = self.stack[i]
            if node.last is not None and node.last[0] == key[i]:
                i += 1
                common = min(node.last[1], out)
                add_prefix = node.last[1] - common
                out -= common
                node.last[1] = common
                if add_prefix:
                    self.stack[i].add_output_prefix(add_prefix)
            else:
                break
        self.len += 1
        self._compile_from(i)
        suffix = key[i:]
        top = self.stack[-1]
        assert top.last is None
        top.last = [suffix[0], out]
        for b in suffix[1:]:
            n = _BNode()
            n.last = [b, 0]
            self.stack.append(n)
        fin = _BNode()
        fin.is_final = True
        self.stack.append(fin)

    def finish(self) -> bytes:
        self._compile_from(0)
        root = self.stack.pop()
        root_addr = self._compile(root)
        self.buf += struct.pack("<QQ", self.len, root_addr)
        return bytes(self.buf)


def build(pairs: Iterable[Tuple[bytes, int]]) -> bytes:
    """buildindex.rs:32-48 semantic: strictly ascending byte keys, u64 values."""
    b = Builder()
    for k, v in pairs:
        b.insert(k, v)
    return b.finish()
