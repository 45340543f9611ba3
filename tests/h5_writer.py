"""Test-side HDF5 *writer* for the subset of the format Keras checkpoints use (tests/test_h5weights.py).

Written from the HDF5 File Format Specification independently of dlwp_cs_b200/h5weights.py (the reader): no HDF5 library
exists in the build image, so reader and writer pin each other.  Two flavours:

  * ``write_earliest``: superblock 0, version-1 object headers (8-byte aligned messages, an attribute message in front of
    the dataset messages, optionally a continuation block), old-style groups (symbol-table message, one level-0 v1 B-tree
    node over SNOD leaves of at most 8 entries, local heap), contiguous datasets -- what h5py/libhdf5 emit by default;
  * ``write_latest``: superblock 2, "OHDR" version-2 headers, compact groups made of link messages, compact datasets.
"""
import struct

import numpy as np

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    big = dt.byteorder == '>'
    if dt.kind == 'f':
        # class 1 (floating point), version 1; bit field: byte order, padding, mantissa normalisation 2 (implied msb), sign position
        size = dt.itemsize
        sign = 8 * size - 1
        bits = (1 if big else 0) | (2 << 4)
        if size == 4:
            props = struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
        else:
            props = struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack('<BBBBI', 0x11, bits, sign, 0, size) + props
    if dt.kind in 'iu':
        bits = (1 if big else 0) | (0x08 if dt.kind == 'i' else 0)
        return struct.pack('<BBBBI', 0x10, bits, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    raise ValueError(dt)


def _dataspace_v1(shape):
    return struct.pack('<BBBBI', 1, len(shape), 0, 0, 0) + b''.join(struct.pack('<Q', s) for s in shape)


def _msg_v1(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHBBBB', mtype, len(body), flags, 0, 0, 0) + body


def _attr_v1(name, value):
    """version-1 attribute message holding a float64 scalar (to be skipped by a weights reader)."""
    nm = name.encode() + b'\x00'
    dt, ds = _dtype_msg('<f8'), struct.pack('<BBBBI', 1, 0, 0, 0, 0)
    return (struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) +
            struct.pack('<d', value))


class _Image(object):
    def __init__(self, start):
        self.buf = bytearray(start)

    def alloc(self, data):
        self.buf += b'\x00' * (-len(self.buf) % 8)
        off = len(self.buf)
        self.buf += data
        return off

    def patch(self, off, data):
        self.buf[off:off + len(data)] = data


def _header_v1(img, msgs, split=False):
    """Object header of the given messages; with `split` the last message goes into a continuation block."""
    if split and len(msgs) > 1:
        tail = msgs[-1]
        cont_off = img.alloc(tail)
        msgs = msgs[:-1] + [_msg_v1(0x10, struct.pack('<QQ', cont_off, len(tail)))]
        n = len(msgs) + 1
    else:
        n = len(msgs)
    body = b''.join(msgs)
    return img.alloc(struct.pack('<BBHII', 1, 0, n, 1, len(body)) + b'\x00' * 4 + body)


def _dataset_v1(img, arr, split=False, with_attr=True):
    arr = np.asarray(arr)
    raw = arr.tobytes()
    data_off = img.alloc(raw)
    msgs = []
    if with_attr:
        msgs.append(_msg_v1(0x0C, _attr_v1('note', 1.5)))
    msgs += [_msg_v1(0x01, _dataspace_v1(arr.shape)), _msg_v1(0x03, _dtype_msg(arr.dtype), flags=1),
             _msg_v1(0x08, struct.pack('<BBQQ', 3, 1, data_off, len(raw)))]
    return _header_v1(img, msgs, split)


def _group_v1(img, children):
    """children: {name: object header offset} -> object header offset of an old-style group."""
    names = sorted(children)
    heap_data = bytearray(b'\x00' * 8)                   # offset 0 = the empty name
    offs = {}
    for nm in names:
        offs[nm] = len(heap_data)
        heap_data += _pad8(nm.encode() + b'\x00')
    heap_data += b'\x00' * 16
    data_off = img.alloc(bytes(heap_data))
    heap_off = img.alloc(b'HEAP' + struct.pack('<BBBB', 0, 0, 0, 0) + struct.pack('<QQQ', len(heap_data), UNDEF, data_off))
    snods, keys = [], [0]
    for i in range(0, max(len(names), 1), 8):
        chunk = names[i:i + 8]
        ent = b''.join(struct.pack('<QQII', offs[nm], children[nm], 0, 0) + b'\x00' * 16 for nm in chunk)
        ent += b'\x00' * (40 * (8 - len(chunk)))
        snods.append(img.alloc(b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk)) + ent))
        keys.append(offs[chunk[-1]] if chunk else 0)
    tree = b'TREE' + struct.pack('<BBH', 0, 0, len(snods)) + struct.pack('<QQ', UNDEF, UNDEF)
    for i, s in enumerate(snods):
        tree += struct.pack('<QQ', keys[i], s)
    tree += struct.pack('<Q', keys[len(snods)])
    tree += b'\x00' * (16 * (32 - len(snods)))          # unused key / child slots of a 2K = 32 entry node
    tree_off = img.alloc(tree)
    return _header_v1(img, [_msg_v1(0x11, struct.pack('<QQ', tree_off, heap_off))]), tree_off, heap_off


def _build_v1(img, node, split):
    if isinstance(node, dict):
        kids = {nm: _build_v1(img, child, split)[0] for nm, child in node.items()}
        return _group_v1(img, kids)
    return _dataset_v1(img, node, split=split), None, None


def write_earliest(path, tree, split_headers=False, userblock=0):
    """tree: nested dict, leaves numpy arrays.  userblock: 0 or 512 (the superblock may sit at offset 512, 1024, ...;
    addresses are then relative to the base address stored in the superblock)."""
    img = _Image(b'\x00' * 96)                            # superblock 0 is 56 + 40 bytes with 8-byte offsets
    root, tree_off, heap_off = _build_v1(img, tree, split_headers)
    sb = SIG + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', 4, 16, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, len(img.buf), UNDEF)
    sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', tree_off, heap_off)
    assert len(sb) == 96
    img.patch(0, sb)
    data = bytes(img.buf)
    if userblock:
        # superblock at `userblock`, base address = userblock: every address in the file is relative to it
        data = bytearray(b'\x00' * userblock + data)
        struct.pack_into('<Q', data, userblock + 24, userblock)
        data = bytes(data)
    with open(path, 'wb') as f:
        f.write(data)


# ---- latest-format flavour ---------------------------------------------------------------------------------------------
def _ohdr(img, msgs):
    body = b''.join(struct.pack('<BHB', t, len(b), 0) + b for t, b in msgs)
    hdr = b'OHDR' + struct.pack('<BB', 2, 0x02) + struct.pack('<I', len(body)) + body     # flags 0x02: 4-byte chunk size
    return img.alloc(hdr + struct.pack('<I', 0))          # checksum not verified by the reader


def _build_latest(img, node):
    if isinstance(node, dict):
        msgs = [(0x02, struct.pack('<BBQQ', 0, 0, UNDEF, UNDEF))]            # link info: no dense storage
        for nm, child in sorted(node.items()):
            off = _build_latest(img, child)
            enc = nm.encode()
            msgs.append((0x06, struct.pack('<BBB', 1, 0x00, len(enc)) + enc + struct.pack('<Q', off)))
        return _ohdr(img, msgs)
    arr = np.asarray(node)
    raw = arr.tobytes()
    ds = struct.pack('<BBBB', 2, arr.ndim, 0, 1) + b''.join(struct.pack('<Q', s) for s in arr.shape)
    return _ohdr(img, [(0x01, ds), (0x03, _dtype_msg(arr.dtype)), (0x08, struct.pack('<BBH', 3, 0, len(raw)) + raw)])


def write_latest(path, tree):
    img = _Image(b'\x00' * 48)
    root = _build_latest(img, tree)
    sb = SIG + struct.pack('<BBBB', 2, 8, 8, 0) + struct.pack('<QQQQ', 0, UNDEF, len(img.buf), root) + struct.pack('<I', 0)
    assert len(sb) == 48
    img.patch(0, sb)
    with open(path, 'wb') as f:
        f.write(bytes(img.buf))
