"""An independent, read-only parser of the HDF5 file-format subset MCRaT's output uses, written
from the HDF5 File Format Specification (superblock v0/v1, v1 object headers, symbol-table groups,
contiguous layout).  TEST INFRASTRUCTURE: it checks the C writer in mcrat_b200/csrc/mcrat_io.c, and
is itself pinned on a file written by the real HDF5 library (tests/test_io.py)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        at = 0
        while self.b[at:at + 8] != SIG:
            at = 512 if at == 0 else at * 2
            if at + 96 > len(self.b):
                raise ValueError("no HDF5 signature")
        sb = self.b[at:]
        self.sb_version = sb[8]
        assert self.sb_version in (0, 1), "superblock version"
        assert sb[9] == 0 and sb[10] == 0 and sb[12] == 0, "free-space / root-entry / shared-header versions"
        assert sb[13] == 8 and sb[14] == 8, "offset / length sizes"
        self.leaf_k, self.int_k = struct.unpack("<HH", sb[16:20])
        v1 = 4 if self.sb_version == 1 else 0
        self.base, self.free, self.eof, self.driver = struct.unpack("<4Q", sb[24 + v1:56 + v1])
        name_off, self.root_ohdr, ctype, _, self.root_btree, self.root_heap = struct.unpack("<QQIIQQ", sb[56 + v1:96 + v1])
        assert name_off == 0 and ctype == 1, "root symbol table entry must cache a group"
        assert self.eof <= len(self.b), "end-of-file address beyond the file"  # libhdf5 stores it inclusive of the base address

    def at(self, addr, n):
        assert addr != UNDEF
        a = addr + self.base
        assert a + n <= len(self.b), "address outside the file"
        return self.b[a:a + n]

    def messages(self, ohdr):
        h = self.at(ohdr, 16)
        ver, _, nmsg, refc, size = struct.unpack("<BBHII", h[:12])
        assert ver == 1 and refc >= 1
        out, blocks = [], [(ohdr + 16, size)]
        while blocks and len(out) < nmsg:
            addr, size = blocks.pop(0)
            p = 0
            data = self.at(addr, size)
            while p + 8 <= size and len(out) < nmsg:
                t, sz, fl = struct.unpack("<HHB", data[p:p + 5])
                assert sz % 8 == 0, "message data must be padded to 8 bytes in a version-1 header"
                body = data[p + 8:p + 8 + sz]
                assert len(body) == sz
                if t == 0x10:
                    blocks.append(struct.unpack("<QQ", body[:16]))
                out.append((t, body))
                p += 8 + sz
        return out

    def group_links(self, btree, heap):
        h = self.at(heap, 32)
        assert h[:4] == b"HEAP" and h[4] == 0
        dsize, free_head, daddr = struct.unpack("<QQQ", h[8:32])
        seg = self.at(daddr, dsize)
        assert seg[0] == 0, "heap offset 0 must hold the empty string"
        # walk the free list: every block inside the segment, terminated by 1 (H5HL_FREE_NULL)
        f, guard = free_head, 0
        while f != 1:
            assert f + 16 <= dsize, "free block outside the heap"
            nxt, size = struct.unpack("<QQ", seg[f:f + 16])
            assert size >= 16 and f + size <= dsize
            f, guard = nxt, guard + 1
            assert guard < 1000
        links = []
        self._walk(btree, seg, links)
        names = [n for n, _, _ in links]
        assert names == sorted(names), "links must be in strcmp order"
        return links

    def _walk(self, node, seg, links):
        t = self.at(node, 24)
        assert t[:4] == b"TREE" and t[4] == 0
        level, used = t[5], struct.unpack("<H", t[6:8])[0]
        assert used <= 2 * self.int_k
        body = self.at(node + 24, (2 * self.int_k + 1) * 8 + 2 * self.int_k * 8)  # the node is allocated at full size
        keys = [struct.unpack("<Q", body[16 * i:16 * i + 8])[0] for i in range(used + 1)]
        kids = [struct.unpack("<Q", body[16 * i + 8:16 * i + 16])[0] for i in range(used)]
        for i, child in enumerate(kids):
            if level > 0:
                self._walk(child, seg, links)
                continue
            s = self.at(child, 8 + 2 * self.leaf_k * 40)  # ditto
            assert s[:4] == b"SNOD" and s[4] == 1
            nsym = struct.unpack("<H", s[6:8])[0]
            assert 1 <= nsym <= 2 * self.leaf_k
            first = last = None
            for k in range(nsym):
                e = s[8 + 40 * k:48 + 40 * k]
                noff, oh, ctype, _ = struct.unpack("<QQII", e[:24])
                name = seg[noff:seg.index(b"\0", noff)].decode()
                scratch = struct.unpack("<QQ", e[24:40]) if ctype == 1 else None
                links.append((name, oh, scratch))
                first = first or name
                last = name
            lo = seg[keys[i]:seg.index(b"\0", keys[i])].decode()
            hi = seg[keys[i + 1]:seg.index(b"\0", keys[i + 1])].decode()
            assert lo < first or (lo == "" and i == 0), "left key must sort before the node's names"
            assert hi == last, "right key must be the node's largest name"

    def object(self, ohdr):
        """-> ('group', {name: ohdr}) or ('dataset', numpy array)"""
        msgs = dict()
        for t, body in self.messages(ohdr):
            msgs.setdefault(t, body)
        if 0x11 in msgs:
            bt, hp = struct.unpack("<QQ", msgs[0x11][:16])
            links = self.group_links(bt, hp)
            for name, oh, scratch in links:
                if scratch is not None:  # cached group info must agree with the child's own header
                    child = dict(self.messages(oh))
                    assert struct.unpack("<QQ", child[0x11][:16]) == scratch
            return "group", {name: oh for name, oh, _ in links}
        sp, ty, lay = msgs[0x01], msgs[0x03], msgs[0x08]
        ver, rank, flags = sp[0], sp[1], sp[2]
        assert ver == 1
        dims = struct.unpack("<%dQ" % rank, sp[8:8 + 8 * rank])
        cls, tver = ty[0] & 0xF, ty[0] >> 4
        size = struct.unpack("<I", ty[4:8])[0]
        assert tver == 1 and (ty[1] & 1) == 0, "little-endian version-1 datatype"
        if cls == 1:
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack("<HHBBBBI", ty[8:20])
            assert (size, off, prec, eloc, esize, mloc, msize, bias) == (8, 0, 64, 52, 11, 0, 52, 1023), "IEEE binary64"
            assert ty[2] == 63 and (ty[1] >> 4) & 3 == 2, "sign bit 63, implied mantissa msb"
            dt = np.dtype("<f8")
        else:
            assert cls == 0
            off, prec = struct.unpack("<HH", ty[8:12])
            assert (size, off, prec) == (1, 0, 8) and (ty[1] & 8), "signed 8-bit integer"
            dt = np.dtype("i1")
        n = int(np.prod(dims)) if rank else 1
        self.last_dataset_info = dict(chunk=None, maxdims=None)
        if flags & 1:
            self.last_dataset_info["maxdims"] = struct.unpack("<%dQ" % rank, sp[8 + 8 * rank:8 + 16 * rank])
        if lay[0] == 3 and lay[1] == 2:
            # chunked storage: dimensionality = rank + 1 (the last one is the element size), version-1 B-tree of node type 1
            ndim = lay[2]
            assert ndim == rank + 1 == 2, "1-D chunked dataset"
            bt = struct.unpack("<Q", lay[3:11])[0]
            cdims = struct.unpack("<%dI" % ndim, lay[11:11 + 4 * ndim])
            assert cdims[-1] == dt.itemsize and cdims[0] > 0
            self.last_dataset_info["chunk"] = cdims[0]
            out = np.zeros(n, dt)
            if bt != UNDEF:
                chunks = self.chunk_tree(bt, ndim, cdims)
                # chunks tile the dataset from 0 in steps of the chunk size, none missing, none past the end
                assert [c[0] for c in chunks] == list(range(0, len(chunks) * cdims[0], cdims[0]))
                assert (len(chunks) - 1) * cdims[0] < n <= len(chunks) * cdims[0]
                for off, nbytes, addr in chunks:
                    assert nbytes == cdims[0] * dt.itemsize
                    part = np.frombuffer(self.at(addr, nbytes), dtype=dt)
                    m = min(cdims[0], n - off)
                    out[off:off + m] = part[:m]
            else:
                assert n == 0
            return "dataset", out.reshape(dims)
        if lay[0] == 3:
            assert lay[1] == 1, "contiguous layout"
            addr, nbytes = struct.unpack("<QQ", lay[2:18])
            assert nbytes == n * dt.itemsize
        else:
            assert lay[0] in (1, 2) and lay[2] == 1
            addr = struct.unpack("<Q", lay[8:16])[0]
        data = np.frombuffer(self.at(addr, n * dt.itemsize), dtype=dt).reshape(dims) if n else np.zeros(dims, dt)
        return "dataset", data

    def chunk_tree(self, node, ndim, cdims, level=None, left=UNDEF):
        """Version-1 B-tree, node type 1 (raw data chunks): -> [(element offset, chunk bytes, address)] in key order.
        Checks the node header, the ascending keys, the key after the last child and the sibling links."""
        keysz = 8 + 8 * ndim
        hdr = self.at(node, 24)
        assert hdr[:4] == b"TREE" and hdr[4] == 1
        lvl, used = hdr[5], struct.unpack("<H", hdr[6:8])[0]
        lsib, rsib = struct.unpack("<QQ", hdr[8:24])
        assert level is None or lvl == level
        assert 0 < used <= 64, "2K entries at most, K = 32 for superblock version 0"
        # the library reads whole nodes: the full-size node must lie inside the file
        self.at(node, 24 + 65 * keysz + 64 * 8)
        body = self.at(node + 24, used * (keysz + 8) + keysz)
        keys, kids = [], []
        for e in range(used + 1):
            k = body[e * (keysz + 8):e * (keysz + 8) + keysz]
            nbytes, mask = struct.unpack("<II", k[:8])
            offs = struct.unpack("<%dQ" % ndim, k[8:])
            keys.append((nbytes, mask, offs))
            if e < used:
                kids.append(struct.unpack("<Q", body[e * (keysz + 8) + keysz:(e + 1) * (keysz + 8)])[0])
        for e in range(used):
            assert keys[e][1] == 0 and keys[e][2][-1] == 0, "no filters; element-dimension offset 0"
            assert keys[e][2][0] % cdims[0] == 0
            assert keys[e][2][0] < keys[e + 1][2][0], "keys ascend; the last key bounds the node from above"
        out = []
        if lvl == 0:
            for e in range(used):
                out.append((keys[e][2][0], keys[e][0], kids[e]))
            self._chunk_leaves.append((node, lsib, rsib))
        else:
            for e in range(used):
                sub = self.chunk_tree(kids[e], ndim, cdims, lvl - 1)
                assert sub[0][0] == keys[e][2][0], "an internal key is the first key of its child"
                out += sub
        return out

    _chunk_leaves = []

    def tree(self):
        """Whole file as nested dicts {name: array | dict}."""
        def rec(oh):
            kind, val = self.object(oh)
            if kind == "dataset":
                return val
            return {k: rec(v) for k, v in val.items()}
        return rec(self.root_ohdr)
