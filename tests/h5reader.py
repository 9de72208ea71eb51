"""An independent reader of the HDF5 subset inovesa_b200/h5native.py writes, straight from the HDF5
File Format Specification: test infrastructure (no HDF5 library exists in this image).  It checks
what a library would check on the way -- signatures, versions, checksums, sizes, addresses inside the
file -- and returns {path: array}, {path: attrs}, {link path: target}."""
import struct

import numpy as np

from inovesa_b200.h5native import lookup3

UNDEF = 0xFFFFFFFFFFFFFFFF


def _dtype(b):
    cls, ver = b[0] & 15, b[0] >> 4
    assert ver == 1, ver
    size = struct.unpack_from("<I", b, 4)[0]
    if cls == 1:
        assert b[1] & 1 == 0 and (b[1] >> 4) & 3 == 2, "little-endian IEEE with implied bit"
        off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", b, 8)
        assert (off, prec, mloc) == (0, 8 * size, 0) and b[2] == 8 * size - 1
        assert (eloc, esize, msize, bias) in ((23, 8, 23, 127), (52, 11, 52, 1023))
        return np.dtype("<f%d" % size), 20
    if cls == 0:
        off, prec = struct.unpack_from("<HH", b, 8)
        assert off == 0 and prec == 8 * size and b[1] & 1 == 0
        return np.dtype("<%s%d" % ("i" if b[1] & 8 else "u", size)), 12
    if cls == 3:
        return np.dtype("S%d" % size), 8
    raise AssertionError("datatype class %d" % cls)


def _space(b):
    assert b[0] == 2, "dataspace version"
    rank, flags, kind = b[1], b[2], b[3]
    assert flags == 0
    if kind == 0:
        assert rank == 0
        return (), 4
    assert kind == 1
    return struct.unpack_from("<%dQ" % rank, b, 4), 4 + 8 * rank


def read_hdf5(path):
    f = open(path, "rb").read()
    assert f[:8] == b"\x89HDF\r\n\x1a\n", "signature"
    assert f[8] == 2 and f[9] == 8 and f[10] == 8, "superblock v2, 8-byte offsets and lengths"
    base, ext, eof, root = struct.unpack_from("<QQQQ", f, 12)
    assert base == 0 and ext == UNDEF and eof == len(f), (base, ext, eof, len(f))
    assert struct.unpack_from("<I", f, 44)[0] == lookup3(f[:44]), "superblock checksum"
    data, attrs, links = {}, {}, {}

    def header(addr):
        assert f[addr:addr + 4] == b"OHDR" and f[addr + 4] == 2, "object header v2 at %d" % addr
        flags = f[addr + 5]
        assert flags & 0x3C == 0 and flags & 3 == 2, "no times / phase change / creation order, 4-byte size"
        size = struct.unpack_from("<I", f, addr + 6)[0]
        end = addr + 10 + size
        assert end + 4 <= len(f)
        assert struct.unpack_from("<I", f, end)[0] == lookup3(f[addr:end]), "object header checksum at %d" % addr
        msgs, p = [], addr + 10
        while p < end:
            mtype, msize, mflags = struct.unpack_from("<BHB", f, p)
            msgs.append((mtype, f[p + 4:p + 4 + msize]))
            p += 4 + msize
        assert p == end, "messages fill the chunk exactly"
        return msgs

    def attribute(b):
        ver, flags, nsz, tsz, ssz, cs = struct.unpack_from("<BBHHHB", b, 0)
        assert ver == 3 and flags == 0
        p = 9
        name = b[p:p + nsz - 1].decode()
        assert b[p + nsz - 1] == 0
        p += nsz
        dt, used = _dtype(b[p:p + tsz])
        assert used == tsz
        p += tsz
        shape, used = _space(b[p:p + ssz])
        assert used == ssz
        p += ssz
        n = int(np.prod(shape)) if shape else 1
        assert len(b) == p + n * dt.itemsize, "attribute data size"
        a = np.frombuffer(b[p:], dt).reshape(shape)
        return name, (a[()] if not shape else a)

    def walk(addr, path):
        kinds = {}
        for mtype, b in header(addr):
            kinds.setdefault(mtype, []).append(b)
        a = dict(attribute(b) for b in kinds.get(0x0C, []))
        if a:
            attrs[path or "/"] = a
        if 0x02 in kinds:  # a group
            ver, flags, heap, btree = struct.unpack_from("<BBQQ", kinds[0x02][0], 0)
            assert (ver, flags, heap, btree) == (0, 0, UNDEF, UNDEF), "compact link storage"
            for b in kinds.get(0x06, []):
                assert b[0] == 1
                lf = b[1]
                assert lf & 0x17 == 0, "1-byte name length, no creation order / charset"
                p = 2
                ltype = 0
                if lf & 8:
                    ltype = b[p]
                    p += 1
                n = b[p]
                name = b[p + 1:p + 1 + n].decode()
                p += 1 + n
                if ltype == 0:
                    child = struct.unpack_from("<Q", b, p)[0]
                    assert p + 8 == len(b) and child < len(f)
                    walk(child, path + "/" + name)
                else:
                    assert ltype == 1
                    tl = struct.unpack_from("<H", b, p)[0]
                    assert p + 2 + tl == len(b)
                    links[path + "/" + name] = b[p + 2:].decode()
            return
        # a dataset
        shape, _ = _space(kinds[0x01][0])
        dt, _ = _dtype(kinds[0x03][0])
        fill = kinds[0x05][0]
        assert fill[0] == 3 and fill[1] & 0x20 == 0
        lay = kinds[0x08][0]
        ver, cls, daddr, dsize = struct.unpack_from("<BBQQ", lay, 0)
        assert ver == 3 and cls == 1, "contiguous layout v3"
        n = int(np.prod(shape)) if shape else 1
        assert dsize == n * dt.itemsize
        if dsize:
            assert daddr % 8 == 0 and daddr + dsize <= len(f)
            data[path] = np.frombuffer(f[daddr:daddr + dsize], dt).reshape(shape)
        else:
            assert daddr == UNDEF
            data[path] = np.zeros(shape, dt)

    walk(root, "")
    return data, attrs, links
