"""Reader for adcc's on-disk SCF data (the ``hfdata`` HDF5 dictionaries of adcc/hdf5io.py:198-224,
keys documented at adcc/DataHfProvider.py:84-141) without h5py.

h5py is not part of this image, so this is a small self-contained parser for the subset of the HDF5
file format those files use (as written by h5py / libhdf5 with default settings): version-0/1
superblock, version-1 object headers with continuation blocks, old-style groups (v1 B-tree + local
heap + symbol nodes), contiguous / compact / chunked (v1 chunk B-tree) dataset layouts, the deflate
and shuffle filters, fixed-point / floating-point / fixed-length string / enum (h5py booleans) /
variable-length string (global heap) datatypes, and version-1 attribute messages.  ``load(fname)``
returns the same nested dict as ``adcc.hdf5io.load`` (datasets tagged by their ``type`` attribute:
scalar / none / ndarray / list / tuple).
"""
import struct
import zlib

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        if buf[:8] != _SIGNATURE:
            raise Hdf5FormatError("not an HDF5 file")
        version = buf[8]
        if version not in (0, 1):
            raise Hdf5FormatError(f"superblock version {version} is not supported")
        self.size_of_offsets, self.size_of_lengths = buf[13], buf[14]
        if (self.size_of_offsets, self.size_of_lengths) != (8, 8):
            raise Hdf5FormatError("only 8-byte offsets / lengths are supported")
        pos = 24 + (4 if version == 1 else 0)
        self.base = self.u64(pos)
        pos += 4 * 8                                  # base, free-space, end-of-file, driver info
        # root group symbol table entry
        self.root_header = self.u64(pos + 8)

    # -- primitive reads
    def u8(self, p): return self.buf[p]
    def u16(self, p): return struct.unpack_from("<H", self.buf, p)[0]
    def u32(self, p): return struct.unpack_from("<I", self.buf, p)[0]
    def u64(self, p): return struct.unpack_from("<Q", self.buf, p)[0]

    # -- object headers
    def messages(self, addr):
        """[(type, payload bytes)] of a version-1 object header including continuation blocks."""
        addr += self.base
        if self.buf[addr] != 1:
            raise Hdf5FormatError("only version-1 object headers are supported")
        n_msgs = self.u16(addr + 2)
        blocks = [(addr + 16, self.u32(addr + 8))]
        out = []
        while blocks and len(out) < n_msgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize = self.u16(pos), self.u16(pos + 2)
                data = self.buf[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:                    # continuation
                    blocks.append((struct.unpack_from("<Q", data, 0)[0] + self.base,
                                   struct.unpack_from("<Q", data, 8)[0]))
                out.append((mtype, data))
        return out

    # -- groups
    def _heap_string(self, heap_addr, offset):
        heap_addr += self.base
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise Hdf5FormatError("local heap signature missing")
        data = self.u64(heap_addr + 24) + self.base
        start = data + offset
        end = self.buf.index(b"\0", start)
        return self.buf[start:end].decode()

    def _group_entries(self, btree, heap):
        """{name: object header address} by walking the group B-tree down to its symbol nodes."""
        out = {}
        node = btree + self.base
        if self.buf[node:node + 4] == b"SNOD":
            n = self.u16(node + 6)
            pos = node + 8
            for _ in range(n):
                out[self._heap_string(heap, self.u64(pos))] = self.u64(pos + 8)
                pos += 40
            return out
        if self.buf[node:node + 4] != b"TREE" or self.buf[node + 4] != 0:
            raise Hdf5FormatError("group B-tree node expected")
        n = self.u16(node + 6)
        pos = node + 24 + 8                          # skip key 0
        for _ in range(n):
            out.update(self._group_entries(self.u64(pos), heap))
            pos += 16                                # child pointer + next key
        return out

    def children(self, header):
        for mtype, data in self.messages(header):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                return self._group_entries(btree, heap)
        return None                                   # not a group

    # -- datatypes / dataspaces
    def parse_datatype(self, data):
        cls, version = data[0] & 0x0F, data[0] >> 4
        bits = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}"), None
        if cls == 1:
            return np.dtype(f"{order}f{size}"), None
        if cls == 3:
            return np.dtype(f"S{size}"), "string"
        if cls == 8:                                  # enum (h5py stores numpy bools like this)
            base, _ = self.parse_datatype(data[8:])
            return base, "enum"
        if cls == 9:
            if bits & 0x0F != 1:
                raise Hdf5FormatError("only variable-length strings are supported")
            return np.dtype("V16"), "vlen_string"
        raise Hdf5FormatError(f"datatype class {cls} (version {version}) is not supported")

    @staticmethod
    def parse_dataspace(data):
        version, rank = data[0], data[1]
        if version == 1:
            pos = 8
        elif version == 2:
            if data[3] == 2:                          # null dataspace
                return None
            pos = 4
        else:
            raise Hdf5FormatError(f"dataspace version {version} is not supported")
        return tuple(struct.unpack_from("<Q", data, pos + 8 * k)[0] for k in range(rank))

    def _vlen_strings(self, raw, count):
        out = []
        for k in range(count):
            length, addr, index = struct.unpack_from("<IQI", raw, 16 * k)
            col = addr + self.base
            if self.buf[col:col + 4] != b"GCOL":
                raise Hdf5FormatError("global heap collection expected")
            end = col + self.u64(col + 8)
            pos = col + 16
            found = None
            while pos + 16 <= end:
                idx, osize = self.u16(pos), self.u64(pos + 8)
                if idx == 0:
                    break
                if idx == index:
                    found = self.buf[pos + 16:pos + 16 + length]
                    break
                pos += 16 + ((osize + 7) // 8) * 8
            if found is None:
                raise Hdf5FormatError("global heap object not found")
            out.append(bytes(found).decode())
        return out

    def _decode(self, raw, dtype, kind, shape):
        count = int(np.prod(shape)) if shape else 1
        if kind == "vlen_string":
            vals = self._vlen_strings(raw, count)
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw, dtype=dtype, count=count)
        if kind == "string":
            vals = [v.split(b"\0")[0].decode() for v in arr.tolist()]
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = arr.astype(dtype.newbyteorder("="))
        if kind == "enum" and dtype.itemsize == 1:
            arr = arr.astype(bool)
        return arr.reshape(shape) if shape else arr.reshape(())[()]

    # -- datasets
    def _chunks(self, node, rank):
        """[(offsets, size, filter mask, address)] of a v1 chunk B-tree."""
        node += self.base
        if self.buf[node:node + 4] != b"TREE" or self.buf[node + 4] != 1:
            raise Hdf5FormatError("chunk B-tree node expected")
        level, n = self.buf[node + 5], self.u16(node + 6)
        key_size = 8 + 8 * (rank + 1)
        pos = node + 24
        out = []
        for _ in range(n):
            size, mask = self.u32(pos), self.u32(pos + 4)
            offs = tuple(self.u64(pos + 8 + 8 * k) for k in range(rank))
            child = self.u64(pos + key_size)
            if level == 0:
                out.append((offs, size, mask, child))
            else:
                out.extend(self._chunks(child, rank))
            pos += key_size + 8
        return out

    def read_dataset(self, msgs):
        shape = dtype = kind = layout = None
        filters = []
        for mtype, data in msgs:
            if mtype == 0x01:
                shape = self.parse_dataspace(data)
            elif mtype == 0x03:
                dtype, kind = self.parse_datatype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._parse_filters(data)
        if dtype is None or layout is None:
            raise Hdf5FormatError("dataset without datatype / layout message")
        if shape is None:
            return None
        if layout[0] != 3:
            raise Hdf5FormatError(f"data layout message version {layout[0]} is not supported")
        lclass = layout[1]
        nbytes = int(np.prod(shape)) * dtype.itemsize if shape else dtype.itemsize
        if lclass == 0:                               # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            raw = layout[4:4 + size]
        elif lclass == 1:                             # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            raw = b"" if addr == _UNDEF else self.buf[addr + self.base:addr + self.base + size]
            if not raw:
                raw = bytes(nbytes)
        elif lclass == 2:                             # chunked
            rank = layout[2] - 1
            btree = struct.unpack_from("<Q", layout, 3)[0]
            cshape = tuple(struct.unpack_from("<I", layout, 11 + 4 * k)[0] for k in range(rank))
            out = np.zeros(shape, dtype=dtype)
            if btree != _UNDEF:
                for offs, size, mask, addr in self._chunks(btree, rank):
                    chunk = bytes(self.buf[addr + self.base:addr + self.base + size])
                    for k, (fid, cd) in reversed(list(enumerate(filters))):
                        if mask & (1 << k):
                            continue
                        if fid == 1:
                            chunk = zlib.decompress(chunk)
                        elif fid == 2:
                            esz = cd[0] if cd else dtype.itemsize
                            a = np.frombuffer(chunk, dtype=np.uint8)
                            chunk = a.reshape(esz, -1).T.tobytes()
                        else:
                            raise Hdf5FormatError(f"filter {fid} is not supported")
                    block = np.frombuffer(chunk, dtype=dtype, count=int(np.prod(cshape))).reshape(cshape)
                    sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                    out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
            raw = out.tobytes()
        else:
            raise Hdf5FormatError(f"layout class {lclass} is not supported")
        return self._decode(raw, dtype, kind, shape)

    @staticmethod
    def _parse_filters(data):
        version, n = data[0], data[1]
        if version != 1:
            raise Hdf5FormatError("only version-1 filter pipeline messages are supported")
        pos = 8
        out = []
        for _ in range(n):
            fid, name_len, _flags, n_cd = struct.unpack_from("<HHHH", data, pos)
            pos += 8 + ((name_len + 7) // 8) * 8
            cd = [struct.unpack_from("<I", data, pos + 4 * k)[0] for k in range(n_cd)]
            pos += 4 * n_cd + (4 if n_cd % 2 else 0)
            out.append((fid, cd))
        return out

    def attributes(self, msgs):
        out = {}
        for mtype, data in msgs:
            if mtype != 0x0C:
                continue
            if data[0] != 1:
                raise Hdf5FormatError("only version-1 attribute messages are supported")
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", data, 2)
            pad = lambda n: ((n + 7) // 8) * 8          # noqa: E731
            pos = 8
            name = bytes(data[pos:pos + name_size]).split(b"\0")[0].decode()
            pos += pad(name_size)
            dtype, kind = self.parse_datatype(data[pos:pos + dt_size])
            pos += pad(dt_size)
            shape = self.parse_dataspace(data[pos:pos + ds_size])
            pos += pad(ds_size)
            out[name] = self._decode(data[pos:], dtype, kind, shape) if shape is not None else None
        return out


def _extract(reader, header):
    kids = reader.children(header)
    if kids is not None:
        return {name: _extract(reader, addr) for name, addr in kids.items()}
    msgs = reader.messages(header)
    value = reader.read_dataset(msgs)
    tpe = reader.attributes(msgs).get("type")
    if tpe is None:
        return value
    if tpe == "none":
        return None
    if tpe == "scalar":
        return value.item() if isinstance(value, np.generic) else value
    if tpe in ("list", "tuple"):
        seq = value.tolist() if isinstance(value, np.ndarray) else list(value)
        return seq if tpe == "list" else tuple(seq)
    if tpe == "ndarray":
        return np.asarray(value)
    raise Hdf5FormatError(f"unknown dataset type tag {tpe!r}")


def load(fname):
    """Nested dict of an adcc HDF5 dictionary file (same result as adcc.hdf5io.load)."""
    with open(fname, "rb") as fp:
        reader = _Reader(fp.read())
    return _extract(reader, reader.root_header)
