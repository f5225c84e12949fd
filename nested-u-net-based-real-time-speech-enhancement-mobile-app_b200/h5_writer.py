"""Dependency-free writer for the classic-HDF5 subset a Keras `save_weights` file uses (the inverse of `h5_reader`).

SURVEY 8(f)4: the step on the other side of the path -- hand the engine's weights back to the reference tool chain as
`nutls_lstm.h5` (`model.load_weights`, `converter_proposed.py:13`).  Neither h5py nor libhdf5 exists in this image, so
the file is assembled byte by byte from the format specification, mirroring the structures found in the reference's own
file: superblock v0, v1 object headers, symbol-table groups (B-tree v1 node + SNOD leaves + local heap), contiguous
little-endian float32 datasets (dataspace v1 with max dims, fill-value message v2, layout v3) and string attributes
(`layer_names`, `weight_names`, `backend`, `keras_version`) as fixed-length strings -- the form older Keras versions
write and h5py returns as byte strings, which `load_attributes_from_hdf5_group` decodes.

What is verified here (tests/test_h5_export.py): the file round-trips through `h5_reader` (datasets and attributes),
reproduces the reference file's group / dataset / attribute inventory, and re-imports to a bit-identical engine blob.
What cannot be verified offline: opening it with libhdf5 itself.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4          # symbol-table leaf: up to 2 K entries per SNOD
INTERNAL_K = 32     # group B-tree node: up to 2 K SNOD children (512 names per group, one node)

AttrValue = Union[str, List[str]]


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _File:
    def __init__(self):
        self.b = bytearray(96)        # superblock, patched at the end

    def alloc(self, n: int) -> int:
        addr = _pad8(len(self.b))
        self.b.extend(b"\0" * (addr - len(self.b) + n))
        return addr

    def put(self, addr: int, data: bytes) -> None:
        self.b[addr:addr + len(data)] = data


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(f: _File, messages: List[bytes]) -> int:
    blob = b"".join(messages)
    addr = f.alloc(16 + len(blob))
    f.put(addr, struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob)
    return addr


def _attr_message(name: str, value: AttrValue) -> bytes:
    vals = [value] if isinstance(value, str) else list(value)
    enc = [v.encode() for v in vals]
    esize = max([len(e) for e in enc] + [1])
    nm = name.encode() + b"\0"
    dtype = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, esize)                 # string, v1, null-padded, ASCII
    if isinstance(value, str):
        dspace = struct.pack("<BBB5x", 1, 0, 0)                           # scalar
    else:
        dspace = struct.pack("<BBB5xQ", 1, 1, 0, len(vals))               # rank 1
    data = b"".join(e + b"\0" * (esize - len(e)) for e in enc)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype), len(dspace))
    body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
    body += dtype + b"\0" * (_pad8(len(dtype)) - len(dtype))
    body += dspace + b"\0" * (_pad8(len(dspace)) - len(dspace))
    return _message(0x0C, body + data)


def _dataset(f: _File, arr: np.ndarray) -> int:
    a = np.ascontiguousarray(arr, dtype="<f4")
    raw = a.tobytes()
    daddr = f.alloc(len(raw)) if raw else UNDEF
    if raw:
        f.put(daddr, raw)
    rank = a.ndim
    dims = struct.pack("<" + "Q" * rank, *a.shape)
    dspace = struct.pack("<BBB5x", 1, rank, 1) + dims + dims                          # v1, max dims = dims
    dtype = struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)   # IEEE float32, little-endian
    fill = bytes([2, 2, 2, 1, 0, 0, 0, 0])                                            # v2, alloc late, never written, size 0
    layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))                              # v3 contiguous
    return _object_header(f, [_message(0x01, dspace), _message(0x03, dtype, 1), _message(0x05, fill, 1), _message(0x08, layout)])


def _group(f: _File, children: Dict[str, tuple], attrs: Dict[str, AttrValue]) -> tuple:
    """children: name -> (header address, cache type, scratch bytes).  Returns (header, btree, heap) addresses."""
    names = sorted(children, key=lambda s: s.encode())
    if len(names) > 2 * LEAF_K * 2 * INTERNAL_K:
        raise ValueError("too many links in one group for a single B-tree node")
    # local heap: "" at offset 0, then the names
    heap_data = bytearray(8)
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        e = n.encode() + b"\0"
        heap_data.extend(e + b"\0" * (_pad8(len(e)) - len(e)))
    heap_data.extend(b"\0" * 16)                       # room libhdf5 likes to find; kept out of the free list
    data_addr = f.alloc(len(heap_data))
    f.put(data_addr, bytes(heap_data))
    heap = f.alloc(32)
    f.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))     # free-list head 1 = none
    # symbol nodes
    snods, keys = [], [0]
    for i in range(0, len(names), 2 * LEAF_K):
        part = names[i:i + 2 * LEAF_K]
        addr = f.alloc(8 + 2 * LEAF_K * 40)
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for n in part:
            hdr, cache, scratch = children[n]
            body += struct.pack("<QQII", offs[n], hdr, cache, 0) + scratch.ljust(16, b"\0")
        f.put(addr, body)
        snods.append(addr)
        keys.append(offs[part[-1]])
    btree = f.alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", keys[0])
    for s, k in zip(snods, keys[1:]):
        node += struct.pack("<QQ", s, k)
    f.put(btree, node)
    msgs = [_message(0x11, struct.pack("<QQ", btree, heap))] + [_attr_message(k, v) for k, v in attrs.items()]
    return _object_header(f, msgs), btree, heap


def write_h5(path: str, datasets: Dict[str, np.ndarray], attrs: Dict[str, Dict[str, AttrValue]]) -> None:
    """`datasets`: '/group/.../name' -> float array; `attrs`: group path ('' = root) -> {name: str | [str]}.  Every
    group named in `attrs` exists in the file even when it has no children (Keras layers without weights)."""
    tree: dict = {}

    def node(path_parts):
        cur = tree
        for p in path_parts:
            cur = cur.setdefault(p, {})
        return cur

    for gp in attrs:
        node([p for p in gp.split("/") if p])
    for key, arr in datasets.items():
        parts = [p for p in key.split("/") if p]
        node(parts[:-1])[parts[-1]] = np.asarray(arr)

    f = _File()

    def emit(sub: dict, prefix: str):
        children = {}
        for name, val in sub.items():
            if isinstance(val, dict):
                hdr, bt, hp = emit(val, prefix + "/" + name)
                children[name] = (hdr, 1, struct.pack("<QQ", bt, hp))
            else:
                children[name] = (_dataset(f, val), 0, b"")
        return _group(f, children, attrs.get(prefix, {}))

    root, bt, hp = emit(tree, "")
    eof = _pad8(len(f.b))
    f.b.extend(b"\0" * (eof - len(f.b)))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, hp)
    assert len(sb) == 96
    f.put(0, sb)
    with open(path, "wb") as fh:
        fh.write(bytes(f.b))
