"""
Minimal pure-Python reader for the HDF5 files Keras writes (``model.save`` / ``save_weights`` -- what the reference's
``DLWP.util.save_model`` produces as ``<name>.keras``, util.py:127-150), so that a reference checkpoint can be loaded into
the engine without h5py (absent from this image) -- SURVEY.md section 8 f4.

Scope -- the subset of the HDF5 File Format Specification (version 3.0) that h5py / libhdf5 emit with their default
("earliest") settings, which is what Keras 2.x / TF 2.1 use:
  * superblock versions 0-3;
  * version-1 object headers (with continuation blocks) and version-2 headers ("OHDR" / "OCHK");
  * old-style groups (symbol-table message -> v1 B-tree "TREE" + "SNOD" leaves + local heap "HEAP") and compact
    new-style groups (link messages); dense link storage (fractal heaps) is rejected with a clear message;
  * datasets with contiguous or compact layout, fixed-point or IEEE floating-point elements of either byte order;
    chunked / filtered storage is rejected (Keras does not chunk its weights).
Attributes are not needed: the tree is walked and every dataset is returned under its path.

Only parsing happens here; nothing is executed from the file.  NOTE on validation: no HDF5 library exists in the build
image, so the reader is pinned against an independent in-repo writer of the same subset (tests/h5_writer.py), both written
from the specification -- see tests/test_h5weights.py.
"""
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _File(object):
    def __init__(self, data):
        self.d = data
        self.base = 0
        self.O = self.L = 8
        self._parse_superblock()

    # ---- primitives --------------------------------------------------------------------------------------------------
    def u(self, off, n):
        if off < 0 or off + n > len(self.d):
            raise H5FormatError('read of %d bytes at %d beyond the end of the file (%d bytes)' % (n, off, len(self.d)))
        return int.from_bytes(self.d[off:off + n], 'little')

    def addr(self, off):
        v = self.u(off, self.O)
        return None if v == (1 << (8 * self.O)) - 1 else v + self.base

    def _parse_superblock(self):
        off = 0
        while True:
            if self.d[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.d):
                raise H5FormatError('not an HDF5 file (signature not found)')
        ver = self.d[off + 8]
        if ver in (0, 1):
            self.O, self.L = self.d[off + 13], self.d[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self.u(p, self.O)
            p += 4 * self.O                                   # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self.root = self.u(p + self.O, self.O) + self.base
        elif ver in (2, 3):
            self.O, self.L = self.d[off + 9], self.d[off + 10]
            p = off + 12
            self.base = self.u(p, self.O)
            self.root = self.u(p + 3 * self.O, self.O) + self.base
        else:
            raise H5FormatError('unsupported superblock version %d' % ver)
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5FormatError('unsupported offset / length sizes %d / %d' % (self.O, self.L))

    # ---- object headers ----------------------------------------------------------------------------------------------
    def messages(self, off):
        """-> list of (type, flags, payload offset, payload size) of the object header at `off`."""
        if self.d[off:off + 4] == b'OHDR':
            return self._messages_v2(off)
        ver = self.d[off]
        if ver != 1:
            raise H5FormatError('unsupported object header version %d at %d' % (ver, off))
        nmsg = self.u(off + 2, 2)
        size = self.u(off + 8, 4)
        out, blocks = [], [(off + 16, size)]
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u(p, 2), self.u(p + 2, 2), self.d[p + 4]
                body = p + 8
                if mtype == 0x10:                              # continuation: offset, length
                    blocks.append((self.addr(body), self.u(body + self.O, self.L)))
                out.append((mtype, mflags, body, msize))
                p = body + msize
        return out

    def _messages_v2(self, off):
        flags = self.d[off + 5]
        p = off + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        n = 1 << (flags & 3)
        size = self.u(p, n)
        p += n
        out, blocks = [], [(p, size)]
        track = bool(flags & 0x04)
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 <= end:
                mtype, msize, mflags = self.d[p], self.u(p + 1, 2), self.d[p + 3]
                body = p + 4 + (2 if track else 0)
                if body + msize > end:
                    break
                if mtype == 0x10:
                    a, ln = self.addr(body), self.u(body + self.O, self.L)
                    if self.d[a:a + 4] != b'OCHK':
                        raise H5FormatError('continuation block without OCHK signature at %d' % a)
                    blocks.append((a + 4, ln - 8))             # minus signature and checksum
                out.append((mtype, mflags, body, msize))
                p = body + msize
        return out

    # ---- groups ------------------------------------------------------------------------------------------------------
    def links(self, off):
        """-> {name: object header address} of the group whose object header is at `off` (None if it is not a group)."""
        msgs = self.messages(off)
        out, is_group = {}, False
        for mtype, _, body, size in msgs:
            if mtype == 0x11:                                   # symbol table: B-tree address, local heap address
                is_group = True
                self._walk_btree(self.addr(body), self._heap_data(self.addr(body + self.O)), out)
            elif mtype == 0x02:                                 # link info: dense storage when the fractal heap address is set
                is_group = True
                lflags = self.d[body + 1]
                p = body + 2 + (8 if lflags & 1 else 0)
                if self.addr(p) is not None:
                    raise H5FormatError('group with dense link storage (fractal heap) is not supported; re-save the model '
                                        'with the default libver')
            elif mtype == 0x06:
                is_group = True
                name, a = self._link_message(body)
                if a is not None:
                    out[name] = a
        return out if is_group else None

    def _heap_data(self, off):
        if self.d[off:off + 4] != b'HEAP':
            raise H5FormatError('local heap signature missing at %d' % off)
        return self.addr(off + 8 + 2 * self.L)

    def _cstr(self, off):
        end = self.d.index(b'\x00', off)
        return self.d[off:end].decode('utf-8')

    def _walk_btree(self, off, heap, out):
        if self.d[off:off + 4] != b'TREE':
            raise H5FormatError('B-tree signature missing at %d' % off)
        if self.d[off + 4] != 0:
            raise H5FormatError('expected a group B-tree node at %d' % off)
        level, used = self.d[off + 5], self.u(off + 6, 2)
        p = off + 8 + 2 * self.O
        for i in range(used):
            child = self.addr(p + self.L + i * (self.L + self.O))
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                if self.d[child:child + 4] != b'SNOD':
                    raise H5FormatError('symbol table node signature missing at %d' % child)
                n = self.u(child + 6, 2)
                q = child + 8
                for _ in range(n):
                    out[self._cstr(heap + self.u(q, self.O))] = self.u(q + self.O, self.O) + self.base
                    q += 2 * self.O + 24

    def _link_message(self, body):
        ver, flags = self.d[body], self.d[body + 1]
        if ver != 1:
            raise H5FormatError('unsupported link message version %d' % ver)
        p = body + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.d[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        n = 1 << (flags & 3)
        ln = self.u(p, n)
        p += n
        name = self.d[p:p + ln].decode('utf-8')
        p += ln
        return name, (self.addr(p) if ltype == 0 else None)    # soft / external links are skipped

    # ---- datasets ----------------------------------------------------------------------------------------------------
    def dataset(self, off):
        """numpy array of the dataset whose object header is at `off`, or None if the object is not a dataset."""
        shape = dtype = layout = None
        for mtype, _, body, size in self.messages(off):
            if mtype == 0x01:
                ver, rank = self.d[body], self.d[body + 1]
                if ver == 1:
                    p = body + 8
                elif ver == 2:
                    p = body + 4
                    if self.d[body + 3] == 2:                  # null dataspace
                        rank = 0
                else:
                    raise H5FormatError('unsupported dataspace version %d' % ver)
                shape = tuple(self.u(p + i * self.L, self.L) for i in range(rank))
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                layout = self._layout(body)
            elif mtype == 0x0B:
                raise H5FormatError('filtered (compressed) datasets are not supported')
        if shape is None or dtype is None or layout is None:
            return None
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        kind, a, b = layout
        if kind == 'contiguous':
            if a is None:                                       # never written: fill value zero
                return np.zeros(shape, dtype=dtype.newbyteorder('='))
            raw = self.d[a:a + count * dtype.itemsize]
        else:
            raw = self.d[a:a + b]
        if len(raw) < count * dtype.itemsize:
            raise H5FormatError('dataset data truncated (%d of %d bytes)' % (len(raw), count * dtype.itemsize))
        arr = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
        return arr.astype(dtype.newbyteorder('='), copy=True)

    def _dtype(self, body):
        cls, bits0, size = self.d[body] & 0x0F, self.d[body + 1], self.u(body + 4, 4)
        order = '>' if bits0 & 1 else '<'
        if cls == 1 and size in (2, 4, 8):
            return np.dtype('%sf%d' % (order, size))
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype('%s%s%d' % (order, 'i' if bits0 & 0x08 else 'u', size))
        raise H5FormatError('unsupported datatype class %d of %d bytes (only fixed- and floating-point numbers)' % (cls, size))

    def _layout(self, body):
        ver = self.d[body]
        if ver == 3:
            cls = self.d[body + 1]
            if cls == 0:
                return ('compact', body + 4, self.u(body + 2, 2))
            if cls == 1:
                return ('contiguous', self.addr(body + 2), self.u(body + 2 + self.O, self.L))
            raise H5FormatError('chunked dataset layout is not supported (Keras stores its weights contiguously)')
        if ver in (1, 2):
            rank, cls = self.d[body + 1], self.d[body + 2]
            p = body + 8
            if cls == 1:
                return ('contiguous', self.addr(p), 0)
            if cls == 0:
                p += 4 * rank
                return ('compact', p + 4, self.u(p, 4))
            raise H5FormatError('chunked dataset layout is not supported')
        raise H5FormatError('unsupported data layout message version %d' % ver)

    def walk(self):
        """-> {path: array} of every dataset in the file."""
        out, seen = {}, set()

        def rec(off, path):
            if off in seen:
                return
            seen.add(off)
            links = self.links(off)
            if links is None:
                arr = self.dataset(off)
                if arr is not None:
                    out[path] = arr
                return
            for name, a in sorted(links.items()):
                rec(a, path + '/' + name if path else name)
        rec(self.root, '')
        return out


def read_datasets(path):
    """{hdf5 path: numpy array} of every numeric dataset in the file."""
    with open(path, 'rb') as f:
        return _File(f.read()).walk()


def read_keras_weights(path):
    """{keras layer name: {weight name: array}} from a Keras HDF5 model / weights file: datasets live at
    ``[model_weights/]<layer>/<layer>/<weight>:0``; the optimizer state (``optimizer_weights``) is ignored."""
    table = {}
    for p, arr in read_datasets(path).items():
        parts = p.split('/')
        if parts[0] == 'optimizer_weights':
            continue
        if parts[0] == 'model_weights':
            parts = parts[1:]
        if len(parts) < 2:
            continue
        weight = parts[-1].rsplit(':', 1)[0]
        table.setdefault(parts[0], {})[weight] = arr
    if not table:
        raise H5FormatError('no layer weights found in %s' % path)
    return table
