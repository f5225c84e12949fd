"""Dependency-free reader for the classic-HDF5 weight file Keras `save_weights` writes.

The reference stores its float32 checkpoint as `dnn_model/log/saved_model/nutls_lstm.h5`
(read by `converter_proposed.py:13` / `test_interface.py:45` through `model.load_weights`).
Neither h5py nor libhdf5 exists in this image, so this module decodes exactly the subset of the
format that file uses: superblock v0, v1 object headers, symbol-table groups (B-tree v1 + SNOD +
local heap) and contiguous little-endian IEEE-float datasets.  Anything else raises.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5FormatError(ValueError):
    pass


class _H5:
    def __init__(self, buf: bytes):
        if buf[:8] != _SIG:
            raise H5FormatError("not an HDF5 file")
        if buf[8] != 0:
            raise H5FormatError("only superblock v0 is supported")
        if buf[13] != 8 or buf[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths are supported")
        self.b = buf
        # root symbol-table entry at byte 56: name off, object header addr, cache type, rsvd, scratch
        self.root_header = struct.unpack_from("<Q", buf, 64)[0]

    # -- object headers ------------------------------------------------------------------------
    def messages(self, addr: int):
        b = self.b
        ver, _, nmsgs, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5FormatError(f"object header v{ver} unsupported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", b, body)
                    blocks.append((off, ln))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- groups --------------------------------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap")
        return struct.unpack_from("<Q", self.b, heap_addr + 24)[0]

    def _name(self, seg: int, off: int) -> str:
        end = self.b.index(b"\x00", seg + off)
        return self.b[seg + off:end].decode()

    def _walk_btree(self, node: int, seg: int, out: Dict[str, int]):
        b = self.b
        if b[node:node + 4] != b"TREE":
            raise H5FormatError("bad group B-tree node")
        _ntype, level, nent = struct.unpack_from("<BBH", b, node + 4)
        pos = node + 24  # after left/right sibling
        for i in range(nent):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level > 0:
                self._walk_btree(child, seg, out)
                continue
            if b[child:child + 4] != b"SNOD":
                raise H5FormatError("bad symbol node")
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            for s in range(nsym):
                name_off, hdr = struct.unpack_from("<QQ", b, child + 8 + s * 40)
                out[self._name(seg, name_off)] = hdr

    def children(self, header: int) -> Dict[str, int] | None:
        for mtype, body, _ in self.messages(header):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.b, body)
                out: Dict[str, int] = {}
                self._walk_btree(btree, self._heap_data(heap), out)
                return out
        return None

    # -- datasets ------------------------------------------------------------------------------
    def dataset(self, header: int) -> np.ndarray:
        b = self.b
        shape: Tuple[int, ...] = ()
        dtype = None
        addr = size = None
        for mtype, body, _ in self.messages(header):
            if mtype == 0x01:
                ver, rank = b[body], b[body + 1]
                if ver != 1:
                    raise H5FormatError("dataspace version")
                shape = struct.unpack_from("<" + "Q" * rank, b, body + 8)
            elif mtype == 0x03:
                cls = b[body] & 0x0F
                sz = struct.unpack_from("<I", b, body + 4)[0]
                if cls == 1 and sz == 4:
                    dtype = np.dtype("<f4")
                elif cls == 1 and sz == 8:
                    dtype = np.dtype("<f8")
                else:
                    raise H5FormatError(f"unsupported datatype class {cls} size {sz}")
            elif mtype == 0x08:
                ver, lclass = b[body], b[body + 1]
                if ver != 3 or lclass != 1:
                    raise H5FormatError("only contiguous v3 layout is supported")
                addr, size = struct.unpack_from("<QQ", b, body + 2)
        if dtype is None or addr is None:
            raise H5FormatError("incomplete dataset header")
        n = int(np.prod(shape)) if shape else 1
        if n * dtype.itemsize != size:
            raise H5FormatError("layout size mismatch")
        return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()


def _attr_strings(h: _H5, body: int, msize: int):
    """Decode one attribute message (v1) holding a string or an array of strings: fixed-length (class 3) or
    variable-length (class 9, elements are global-heap references) -- the two forms h5py writes.  Returns (name, value)."""
    b = h.b
    if b[body] != 1:
        raise H5FormatError("attribute message version")
    nsz, dtsz, dssz = struct.unpack_from("<HHH", b, body + 2)
    pad = lambda x: (x + 7) // 8 * 8
    name = b[body + 8:body + 8 + nsz].split(b"\0")[0].decode()
    dt = body + 8 + pad(nsz)
    ds = dt + pad(dtsz)
    data = ds + pad(dssz)
    rank = b[ds + 1]
    dims = struct.unpack_from("<" + "Q" * rank, b, ds + 8) if rank else ()
    n = int(np.prod(dims)) if rank else 1
    cls = b[dt] & 0x0F
    esize = struct.unpack_from("<I", b, dt + 4)[0]
    vals = []
    for i in range(n):
        if cls == 3:
            raw = b[data + i * esize:data + (i + 1) * esize]
            vals.append(raw.split(b"\0")[0].decode())
        elif cls == 9:
            ln, addr, idx = struct.unpack_from("<IQI", b, data + i * 16)
            if b[addr:addr + 4] != b"GCOL":
                raise H5FormatError("bad global heap")
            pos, s_val = addr + 16, None
            end = addr + struct.unpack_from("<Q", b, addr + 8)[0]
            while pos + 16 <= end:
                oidx, _ref, _r, osize = struct.unpack_from("<HHIQ", b, pos)
                if oidx == idx:
                    s_val = b[pos + 16:pos + 16 + ln].decode()
                    break
                if oidx == 0:
                    break
                pos += 16 + pad(osize)
            if s_val is None:
                raise H5FormatError("global heap object not found")
            vals.append(s_val)
        else:
            raise H5FormatError(f"unsupported attribute datatype class {cls}")
    return name, (vals if rank else vals[0])


def read_h5_attrs(path: str) -> Dict[str, Dict[str, object]]:
    """String attributes of every group: `{'': {'layer_names': [...], 'backend': ..}, '/conv2d': {'weight_names': [...]}}`
    (what `keras.Model.load_weights` walks: `layer_names` on the root, `weight_names` on each layer group)."""
    with open(path, "rb") as f:
        h = _H5(f.read())
    out: Dict[str, Dict[str, object]] = {}

    def rec(header: int, prefix: str):
        kids = h.children(header)
        if kids is None:
            return
        attrs = {}
        for mtype, body, msize in h.messages(header):
            if mtype == 0x0C:
                k, v = _attr_strings(h, body, msize)
                attrs[k] = v
        out[prefix] = attrs
        for name, hdr in kids.items():
            rec(hdr, prefix + "/" + name)

    rec(h.root_header, "")
    return out


def read_h5(path: str) -> Dict[str, np.ndarray]:
    """Return `{'/group/.../name:0': ndarray}` for every dataset in a Keras weight file."""
    with open(path, "rb") as f:
        h = _H5(f.read())
    out: Dict[str, np.ndarray] = {}

    def rec(header: int, prefix: str):
        kids = h.children(header)
        if kids is None:
            out[prefix] = h.dataset(header)
            return
        for name, hdr in kids.items():
            rec(hdr, prefix + "/" + name)

    rec(h.root_header, "")
    return out
