"""Minimal read-only HDF5 reader (h5py is not available in this image).

Covers exactly the subset the reference's artefacts use:

* ``Weights.h5`` written by Keras 2.7 ``model.save_weights`` (cavity_steady.py:249-252):
  superblock v0, old-style groups (symbol table + B-tree v1 + local heap), object headers v1,
  contiguous little-endian float datasets, fixed-length / variable-length string attributes are
  skipped (the layer order is recovered from the group tree);
* FEniCS XDMF companions (``VisualisationVector/0``, ``/1``, ``Mesh/0/mesh/geometry``;
  cavity_steady.py:100-105) -- same structural subset, contiguous float64/int datasets.

Not supported (raises): chunked/compressed layouts, new-style (v2) groups, superblock >= 2.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class _Dataset:
    def __init__(self, shape, dtype, addr, size):
        self.shape, self.dtype, self.addr, self.size = shape, dtype, addr, size


class H5File:
    """``H5File(path)[ "a/b/c" ]`` -> numpy array; ``.datasets()`` lists every dataset path."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise H5Error("not an HDF5 file")
        ver = self.buf[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} not supported")
        self.off_size = self.buf[13]
        self.len_size = self.buf[14]
        if self.off_size != 8 or self.len_size != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        # v0: sig(8) ver(1) fsver(1) rootver(1) res(1) shmver(1) offsz(1) lensz(1) res(1)
        #     leafk(2) intk(2) flags(4) [v1: indexed storage k(2) res(2)] base(8) free(8) eof(8) drv(8)
        pos = 24 + (4 if ver == 1 else 0)
        pos += 32
        # root symbol table entry: link name offset(8) obj header addr(8) cache type(4) res(4) scratch(16)
        _, self.root_addr, cache_type = struct.unpack_from("<QQI", self.buf, pos)
        self._tree: Dict[str, object] = {}
        self._walk_group(self.root_addr, "")

    # ---- object headers -----------------------------------------------------------------
    def _messages(self, addr: int) -> Iterator[Tuple[int, bytes]]:
        b = self.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr} not supported")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                else:
                    yield mtype, body

    def _walk_group(self, addr: int, prefix: str) -> None:
        stab = None
        ds_space = ds_type = ds_layout = None
        for mtype, body in self._messages(addr):
            if mtype == 0x0011:
                stab = struct.unpack_from("<QQ", body, 0)
            elif mtype == 0x0001:
                ds_space = body
            elif mtype == 0x0003:
                ds_type = body
            elif mtype == 0x0008:
                ds_layout = body
        if stab is not None:
            btree, heap = stab
            heap_data = self._local_heap(heap)
            for name_off, obj_addr in self._btree_entries(btree):
                end = self.buf.index(b"\x00", heap_data + name_off)
                name = self.buf[heap_data + name_off:end].decode()
                self._walk_group(obj_addr, f"{prefix}/{name}" if prefix else name)
        elif ds_space is not None and ds_type is not None and ds_layout is not None:
            self._tree[prefix] = self._dataset(ds_space, ds_type, ds_layout)

    def _local_heap(self, addr: int) -> int:
        if self.buf[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        _size, _free, data = struct.unpack_from("<QQQ", self.buf, addr + 8)
        return data

    def _btree_entries(self, addr: int) -> List[Tuple[int, int]]:
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        pos = addr + 8 + 16  # skip siblings
        out: List[Tuple[int, int]] = []
        pos += 8  # key 0
        for _ in range(used):
            child = struct.unpack_from("<Q", b, pos)[0]
            pos += 16  # child + next key
            if level > 0:
                out += self._btree_entries(child)
            else:
                out += self._snod(child)
        return out

    def _snod(self, addr: int) -> List[Tuple[int, int]]:
        b = self.buf
        if b[addr:addr + 4] != b"SNOD":
            raise H5Error("bad symbol node signature")
        nsym = struct.unpack_from("<H", b, addr + 6)[0]
        out = []
        pos = addr + 8
        for _ in range(nsym):
            name_off, obj = struct.unpack_from("<QQ", b, pos)
            out.append((name_off, obj))
            pos += 40
        return out

    # ---- datasets -----------------------------------------------------------------------
    def _dataset(self, space: bytes, dtype: bytes, layout: bytes) -> _Dataset:
        sver, rank, sflags = struct.unpack_from("<BBB", space, 0)
        if sver == 1:
            dims = struct.unpack_from(f"<{rank}Q", space, 8)
        elif sver == 2:
            dims = struct.unpack_from(f"<{rank}Q", space, 4)
        else:
            raise H5Error(f"dataspace version {sver}")
        cls = dtype[0] & 0x0F
        bits0 = dtype[1]
        size = struct.unpack_from("<I", dtype, 4)[0]
        if bits0 & 1:
            raise H5Error("big-endian data not supported")
        if cls == 1:
            npdt = {4: "<f4", 8: "<f8"}[size]
        elif cls == 0:
            signed = bool(bits0 & 0x08)
            npdt = ("<i" if signed else "<u") + str(size)
        else:
            raise H5Error(f"datatype class {cls} not supported")
        lver = layout[0]
        if lver == 3:
            lclass = layout[1]
            if lclass != 1:
                raise H5Error("only contiguous layout supported")
            addr, nbytes = struct.unpack_from("<QQ", layout, 2)
        elif lver in (1, 2):
            lrank, lclass = layout[1], layout[2]
            if lclass != 1:
                raise H5Error("only contiguous layout supported")
            addr = struct.unpack_from("<Q", layout, 8)[0]
            nbytes = int(np.prod(dims, dtype=np.int64)) * size
        else:
            raise H5Error(f"layout version {lver}")
        return _Dataset(tuple(int(d) for d in dims), npdt, addr, nbytes)

    # ---- public -------------------------------------------------------------------------
    def datasets(self) -> List[str]:
        return list(self._tree.keys())

    def __contains__(self, key: str) -> bool:
        return key.strip("/") in self._tree

    def __getitem__(self, key: str) -> np.ndarray:
        ds = self._tree[key.strip("/")]
        n = int(np.prod(ds.shape, dtype=np.int64)) if ds.shape else 1
        if ds.addr == _UNDEF:
            return np.zeros(ds.shape, dtype=ds.dtype)
        arr = np.frombuffer(self.buf, dtype=ds.dtype, count=n, offset=ds.addr)
        return arr.reshape(ds.shape).copy()


def load_keras_dense_weights(path: str) -> List[np.ndarray]:
    """Return ``[K1, b1, K2, b2, ...]`` (Keras variable order) from a ``Weights.h5``.

    Layer groups are named ``dense_<n>``; Keras numbers layers in creation order, so sorting by the
    numeric suffix (bare ``dense`` first) reproduces ``model.variables`` order.
    """
    f = H5File(path)
    layers: Dict[str, Dict[str, np.ndarray]] = {}
    for p in f.datasets():
        parts = p.split("/")
        leaf = parts[-1]
        if leaf in ("kernel:0", "bias:0"):
            layers.setdefault(parts[0], {})[leaf] = f[p]

    def key(name: str) -> int:
        tail = name.rsplit("_", 1)
        return int(tail[1]) if len(tail) == 2 and tail[1].isdigit() else -1

    out: List[np.ndarray] = []
    for name in sorted(layers, key=key):
        out += [layers[name]["kernel:0"], layers[name]["bias:0"]]
    return out


# ------------------------------------------------------------------------------------------------
# writer: the same structural subset (superblock v0, old-style groups, contiguous datasets)
# ------------------------------------------------------------------------------------------------

_LEAF_K, _INT_K = 4, 16          # group B-tree parameters recorded in the superblock


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(messages: List[bytes]) -> bytes:
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        size = dt.itemsize
        exp_bits, man_bits, bias = (8, 23, 127) if size == 4 else (11, 52, 1023)
        # class 1 (floating point), version 1; bit field 0: little-endian, mantissa normalisation 2 (implied msb); bit field 1: sign position
        head = struct.pack("<BBBBI", 0x11, 0x20, 8 * size - 1, 0, size)
        return head + struct.pack("<HHBBBBI", 0, 8 * size, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise H5Error(f"dtype {dt} not supported by the writer")


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)      # superblock placeholder

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else _UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        return self.alloc(_object_header([_message(0x0001, space), _message(0x0003, _datatype_message(arr.dtype)),
                                          _message(0x0008, layout)]))

    def group(self, children: Dict[str, int], is_group: Dict[str, Tuple[int, int]]) -> Tuple[int, int, int]:
        """children: name -> object header address.  Returns (header address, B-tree address, heap address)."""
        names = sorted(children)
        if len(names) > 2 * _LEAF_K * 2 * _INT_K:
            raise H5Error("too many links in one group for the single-level B-tree of this writer")
        heap_data = bytearray(8)      # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\x00")
        heap_seg = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _UNDEF, heap_seg))
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 2 * _LEAF_K):
            part = names[i:i + 2 * _LEAF_K]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                if n in is_group:
                    body += struct.pack("<QQII", offs[n], children[n], 1, 0) + struct.pack("<QQ", *is_group[n])
                else:
                    body += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            body += b"\x00" * (8 + 2 * _LEAF_K * 40 - len(body))
            snods.append(self.alloc(body))
            keys.append(offs[part[-1]] if part else 0)
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
        for i, a in enumerate(snods):
            node += struct.pack("<QQ", keys[i], a)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\x00" * (24 + (2 * _INT_K) * 8 + (2 * _INT_K + 1) * 8 - len(node))
        btree = self.alloc(node)
        header = self.alloc(_object_header([_message(0x0011, struct.pack("<QQ", btree, heap))]))
        return header, btree, heap


def write_h5(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write ``{"group/sub/name": array}`` as an HDF5 file of the subset ``H5File`` reads (and h5py / Keras read too):
    superblock v0, old-style groups, contiguous little-endian float / integer datasets, no attributes."""
    tree: Dict[str, object] = {}
    for key, arr in datasets.items():
        parts = [p for p in key.split("/") if p]
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise H5Error(f"{key}: a dataset is in the way")
        node[parts[-1]] = np.asarray(arr)
    w = _Writer()

    def emit(node) -> Tuple[int, int, int]:
        children, groups = {}, {}
        for name, val in node.items():
            if isinstance(val, dict):
                h, b, hp = emit(val)
                children[name], groups[name] = h, (b, hp)
            else:
                children[name] = w.dataset(val)
        return w.group(children, groups)

    root, btree, heap = emit(tree)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INT_K, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, len(w.buf), _UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))


def save_keras_dense_weights(path: str, variables: List[np.ndarray], dtype=np.float64) -> None:
    """``[K1, b1, K2, b2, ...]`` -> a ``Weights.h5`` with the dataset paths Keras 2.7's ``model.save_weights`` uses for a
    Sequential of Dense layers (``dense/dense/kernel:0``, ``dense_1/dense_1/bias:0``, ...; cavity_steady.py:249-252).
    The reference's files also carry string attributes (``layer_names``, ``weight_names``) that this writer does not emit:
    ``load_keras_dense_weights`` and any reader that walks the groups do not need them."""
    out = {}
    for i in range(len(variables) // 2):
        name = "dense" if i == 0 else f"dense_{i}"
        out[f"{name}/{name}/kernel:0"] = np.asarray(variables[2 * i], dtype=dtype)
        out[f"{name}/{name}/bias:0"] = np.asarray(variables[2 * i + 1], dtype=dtype)
    write_h5(path, out)
