"""A minimal HDF5 writer in pure Python + numpy, for result files where neither libhdf5 nor h5py
exist (this image).  It writes the subset of the HDF5 File Format Specification (version 3.0) that
a result file in Inovesa's layout (src/IO/HDF5File.cpp:29-349) needs to be READ by any HDF5 library:

* superblock version 2, object headers version 2 (with their Jenkins lookup3 checksums);
* groups with compact link storage (Link Info + Link messages in the object header): hard links
  to groups and datasets, soft links (the reference's `axis0` / `axis1` links);
* datasets with contiguous storage: little-endian IEEE floats and integers of 1-8 bytes, any rank
  (rank 0 = scalar, empty extents allowed), fill-value message version 3;
* attributes (version 3) on groups and datasets: numeric scalars / arrays and fixed-length strings.

What it does not write: chunked / extensible / compressed storage (the reference's datasets are
chunked, shuffled and deflated; a reader does not see the difference), dense link or attribute
storage, compound types, references.

    write_hdf5(path, datasets={"/A/data": array, ...}, attrs={"/A/data": {"Meter": 1e-3}},
               groups=["/Empty"], soft_links={"/A/axis0": "/Info/AxisValues_t"})
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_M = 0xFFFFFFFF


def lookup3(data, initval=0):
    """Bob Jenkins' hashlittle() (lookup3.c), the checksum of version-2 structures (H5_checksum_lookup3)"""
    def rot(x, k):
        return ((x << k) | (x >> (32 - k))) & _M

    data = bytes(data)
    length = len(data)
    a = b = c = (0xDEADBEEF + length + initval) & _M
    i = 0
    while length > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & _M
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & _M
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & _M
        a = (a - c) & _M; a ^= rot(c, 4); c = (c + b) & _M
        b = (b - a) & _M; b ^= rot(a, 6); a = (a + c) & _M
        c = (c - b) & _M; c ^= rot(b, 8); b = (b + a) & _M
        a = (a - c) & _M; a ^= rot(c, 16); c = (c + b) & _M
        b = (b - a) & _M; b ^= rot(a, 19); a = (a + c) & _M
        c = (c - b) & _M; c ^= rot(b, 4); b = (b + a) & _M
        i += 12
        length -= 12
    if length == 0:
        return c
    tail = data[i:] + b"\0" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & _M
    b = (b + int.from_bytes(tail[4:8], "little")) & _M
    c = (c + int.from_bytes(tail[8:12], "little")) & _M
    c ^= b; c = (c - rot(b, 14)) & _M
    a ^= c; a = (a - rot(c, 11)) & _M
    b ^= a; b = (b - rot(a, 25)) & _M
    c ^= b; c = (c - rot(b, 16)) & _M
    a ^= c; a = (a - rot(c, 4)) & _M
    b ^= a; b = (b - rot(a, 14)) & _M
    c ^= b; c = (c - rot(b, 24)) & _M
    return c


# ---- message bodies --------------------------------------------------------------------------

def _datatype(dt):
    """Datatype message body (IV.A.2.d) of a little-endian numpy dtype"""
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        elif dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        else:
            raise TypeError("float size %d" % dt.itemsize)
        # class 1 (floating point), version 1; little-endian, IEEE implied leading bit (mantissa normalisation 2)
        head = bytes([0x11, 0x20, sign, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
    if dt.kind in "iub":
        signed = 0x08 if dt.kind == "i" else 0x00
        head = bytes([0x10, signed, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        # class 3 (string), null-padded ASCII
        return bytes([0x13, 0x01, 0x00, 0x00]) + struct.pack("<I", max(dt.itemsize, 1))
    raise TypeError("unsupported dtype %s" % dt)


def _dataspace(shape):
    """Dataspace message body, version 2 (IV.A.2.b): scalar or simple, no maximum dimensions"""
    shape = tuple(int(s) for s in shape)
    if not shape:
        return bytes([2, 0, 0, 0])
    return bytes([2, len(shape), 0, 1]) + b"".join(struct.pack("<Q", s) for s in shape)


def _as_array(value):
    if isinstance(value, str):
        value = value.encode("ascii", "replace")
    if isinstance(value, (bytes, bytearray)):
        return np.array(bytes(value) or b"\0", dtype="S%d" % max(len(value), 1))
    a = np.asarray(value)
    if a.dtype.kind == "b":
        a = a.astype(np.uint8)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "ascii")
    if a.dtype.kind not in "fiuS":
        raise TypeError("unsupported value of dtype %s" % a.dtype)
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)
    return a if a.ndim == 0 else np.ascontiguousarray(a)  # (ascontiguousarray would make a scalar 1-D)


def _message(mtype, body):
    """header message of a version-2 object header (no creation order): type, size, flags, body"""
    return struct.pack("<BHB", mtype, len(body), 0) + body


def _attribute(name, value):
    a = _as_array(value)
    nm = name.encode("utf-8") + b"\0"
    dt, ds = _datatype(a.dtype), _dataspace(a.shape)
    body = struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 0) + nm + dt + ds + a.tobytes()
    return _message(0x0C, body)


def _link(name, addr=None, soft=None):
    nm = name.encode("utf-8")
    assert 0 < len(nm) < 256
    if soft is None:
        return _message(0x06, struct.pack("<BBB", 1, 0x00, len(nm)) + nm + struct.pack("<Q", addr))
    tgt = soft.encode("utf-8")
    return _message(0x06, struct.pack("<BBBB", 1, 0x08, 1, len(nm)) + nm + struct.pack("<H", len(tgt)) + tgt)


def _object_header(messages):
    """OHDR, version 2, 4-byte chunk size, no times, no attribute phase change values; checksum"""
    body = b"".join(messages)
    head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body
    return head + struct.pack("<I", lookup3(head))


# ---- the file --------------------------------------------------------------------------------

class _Node:
    def __init__(self):
        self.children = {}   # name -> _Node (group) / ("data", array) / ("soft", target)
        self.attrs = {}
        self.addr = None


def _split(path):
    parts = [p for p in path.split("/") if p]
    if not parts:
        raise ValueError("empty path")
    return parts


def write_hdf5(path, datasets, attrs=None, groups=(), soft_links=None):
    """write one HDF5 file; paths are absolute ('/Group/name'); attrs may name groups or datasets"""
    root = _Node()

    def group(parts):
        node = root
        for p in parts:
            nxt = node.children.get(p)
            if nxt is None:
                nxt = node.children[p] = _Node()
            if not isinstance(nxt, _Node):
                raise ValueError("/%s is not a group" % "/".join(parts))
            node = nxt
        return node

    for g in groups:
        group(_split(g))
    data_attrs = {}
    for p, arr in datasets.items():
        parts = _split(p)
        parent = group(parts[:-1])
        if parts[-1] in parent.children:
            raise ValueError("%s exists twice" % p)
        parent.children[parts[-1]] = ("data", _as_array(arr), p)
    for p, target in (soft_links or {}).items():
        parts = _split(p)
        group(parts[:-1]).children[parts[-1]] = ("soft", target)
    for p, kv in (attrs or {}).items():
        if p in datasets:
            data_attrs[p] = dict(kv)
        else:
            group(_split(p) if p.strip("/") else []).attrs.update(kv)

    # pass 1: sizes (they do not depend on the addresses: every address field is 8 bytes), then addresses
    def dataset_header(arr, p, addr):
        msgs = [_message(0x01, _dataspace(arr.shape)), _message(0x03, _datatype(arr.dtype)),
                _message(0x05, bytes([3, 0x0A])),  # fill value v3: late allocation, written if set, none defined
                _message(0x08, struct.pack("<BBQQ", 3, 1, addr if arr.nbytes else UNDEF, arr.nbytes))]
        msgs += [_attribute(k, v) for k, v in data_attrs.get(p, {}).items()]
        return _object_header(msgs)

    def group_header(node):
        msgs = [_message(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)),  # link info: compact storage
                _message(0x0A, bytes([0, 0]))]                               # group info: defaults
        for name in sorted(node.children):
            ch = node.children[name]
            if isinstance(ch, _Node):
                msgs.append(_link(name, addr=ch.addr or 0))
            elif ch[0] == "data":
                msgs.append(_link(name, addr=ch[3] if len(ch) > 3 else 0))
            else:
                msgs.append(_link(name, soft=ch[1]))
        msgs += [_attribute(k, v) for k, v in node.attrs.items()]
        return _object_header(msgs)

    pos = [48]  # after the superblock

    def place(node):
        node.addr = pos[0]
        pos[0] += len(group_header(node))
        for name in sorted(node.children):
            ch = node.children[name]
            if isinstance(ch, _Node):
                place(ch)
            elif ch[0] == "data":
                _, arr, p = ch[:3]
                haddr = pos[0]
                pos[0] += len(dataset_header(arr, p, 0))
                pos[0] = (pos[0] + 7) & ~7
                daddr = pos[0]
                pos[0] += arr.nbytes
                node.children[name] = ("data", arr, p, haddr, daddr)

    place(root)
    eof = pos[0]

    # pass 2: emit
    out = bytearray(eof)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, root.addr)
    out[0:48] = sb + struct.pack("<I", lookup3(sb))

    def emit(node):
        h = group_header(node)
        out[node.addr:node.addr + len(h)] = h
        for name in sorted(node.children):
            ch = node.children[name]
            if isinstance(ch, _Node):
                emit(ch)
            elif ch[0] == "data":
                _, arr, p, haddr, daddr = ch
                h = dataset_header(arr, p, daddr)
                out[haddr:haddr + len(h)] = h
                out[daddr:daddr + arr.nbytes] = arr.tobytes()

    emit(root)
    with open(path, "wb") as f:
        f.write(out)
    return eof


# ---- reading back what write_hdf5 wrote (start distributions from an earlier result file) -----

def read_dataset(path, dataset):
    """the array of one dataset of a file written by write_hdf5 (superblock v2, compact groups, contiguous
    little-endian data); files of libhdf5 with chunked storage need h5py"""
    f = open(path, "rb").read()
    if f[:8] != b"\x89HDF\r\n\x1a\n":
        raise ValueError("%s is not an HDF5 file" % path)
    if f[8] != 2 or f[9] != 8 or f[10] != 8:
        raise ValueError("%s: superblock version %d is not what this reader knows (use h5py)" % (path, f[8]))
    if struct.unpack_from("<I", f, 44)[0] != lookup3(f[:44]):
        raise ValueError("%s: superblock checksum" % path)
    addr = struct.unpack_from("<Q", f, 36)[0]

    def messages(a):
        if f[a:a + 4] != b"OHDR" or f[a + 4] != 2 or f[a + 5] != 0x02:
            raise ValueError("%s: object header at %d is not of the kind write_hdf5 writes (use h5py)" % (path, a))
        size = struct.unpack_from("<I", f, a + 6)[0]
        end = a + 10 + size
        if struct.unpack_from("<I", f, end)[0] != lookup3(f[a:end]):
            raise ValueError("%s: object header checksum at %d" % (path, a))
        p = a + 10
        while p < end:
            mtype, msize, _ = struct.unpack_from("<BHB", f, p)
            yield mtype, f[p + 4:p + 4 + msize]
            p += 4 + msize

    for name in _split(dataset):
        nxt = None
        for mtype, b in messages(addr):
            if mtype != 0x06 or b[1] & 0x08:  # hard links only
                continue
            n = b[2]
            if b[3:3 + n].decode() == name:
                nxt = struct.unpack_from("<Q", b, 3 + n)[0]
        if nxt is None:
            raise KeyError("%s: no %s" % (path, dataset))
        addr = nxt
    shape = dt = lay = None
    for mtype, b in messages(addr):
        if mtype == 0x01:
            shape = struct.unpack_from("<%dQ" % b[1], b, 4) if b[1] else ()
        elif mtype == 0x03:
            size = struct.unpack_from("<I", b, 4)[0]
            cls = b[0] & 15
            dt = np.dtype("<f%d" % size) if cls == 1 else np.dtype("<%s%d" % ("i" if b[1] & 8 else "u", size)) if cls == 0 else None
        elif mtype == 0x08:
            lay = struct.unpack_from("<BBQQ", b, 0)
    if shape is None or dt is None or lay is None or lay[1] != 1:
        raise ValueError("%s: %s is not a contiguous numeric dataset (use h5py)" % (path, dataset))
    if lay[3] == 0:
        return np.zeros(shape, dt)
    return np.frombuffer(f[lay[2]:lay[2] + lay[3]], dt).reshape(shape).copy()
