"""Minimal HDF5 reader / writer for the feature hand-off files of the reference — no libhdf5, no h5py.

The reference moves per-frame embeddings between its stages through ``<model>_RepsAndLabels.h5`` /
``<model>_FlowRepsAndLabels.h5`` (``saveH5``, extract_representations.py:389-407; read back at
prepare_dataset.py:317-319, 2658-2667): ONE flat root group holding one contiguous float32 ``[n_frames,384]`` dataset per
video label.  That is a small, fixed subset of the HDF5 file format, written and parsed here directly from the format
specification (HDF5 File Format Specification, version 1.1 structures — what libhdf5 / h5py write by default):

* superblock version 0 (8-byte offsets and lengths), optional user block (base address) honoured on read;
* "old style" groups: symbol-table message -> version-1 B-tree (node type 0) -> symbol-table nodes (``SNOD``) ->
  names in a local heap (``HEAP``);
* version-1 object headers (continuation blocks followed on read), dataspace message v1 / v2, datatype message for
  fixed-point and IEEE floating-point classes (little- or big-endian), data-layout message v3 (contiguous and compact;
  v1 / v2 contiguous on read), fill-value message v1.

Not covered (raises :class:`H5LiteError`): chunked / filtered layouts, version-2 ("new style") groups and object headers,
variable-length / compound types, attributes (skipped on read).  ``h5py.File(..., 'w')`` + ``create_dataset(name, data=...)``
with default settings — the only thing the reference does — produces exactly the covered subset.

Validation: the reader is pinned against a file written by libhdf5 itself (the MATLAB 7.3 test file that ships with
SciPy, ``tests/test_postprocess.py::test_h5lite_reads_a_libhdf5_file``); the writer is checked through that reader,
structure by structure against the same file's layout, and byte-for-byte on the raw data segment.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteError(RuntimeError):
    pass


# ======================================================================================================== reader
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        sig = -1
        off = 0
        while off < len(buf):  # the superblock sits at 0, 512, 1024, 2048, ... (a user block may precede it)
            if buf[off:off + 8] == SIGNATURE:
                sig = off
                break
            off = 512 if off == 0 else off * 2
        if sig < 0:
            raise H5LiteError("not an HDF5 file (no superblock signature)")
        ver = buf[sig + 8]
        if ver not in (0, 1):
            raise H5LiteError(f"superblock version {ver} is not supported (only the default version 0 / 1 layout)")
        self.so, self.sl = buf[sig + 13], buf[sig + 14]
        if (self.so, self.sl) != (8, 8):
            raise H5LiteError("only 8-byte offsets / lengths are supported")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", buf, sig + 16)
        p = sig + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        self.root_entry = self._sym_entry(p + 32)

    # -- primitives (addresses in the file are relative to the base address)
    def _at(self, addr: int) -> int:
        return addr + self.base

    def _sym_entry(self, p: int):
        name_off, ohdr, cache = struct.unpack_from("<QQI", self.b, p)
        scratch = self.b[p + 24:p + 40]
        return {"name_off": name_off, "ohdr": ohdr, "cache": cache, "scratch": scratch}

    def _heap(self, addr: int) -> Tuple[int, int]:
        p = self._at(addr)
        if self.b[p:p + 4] != b"HEAP":
            raise H5LiteError("local heap signature missing")
        size, _free, data = struct.unpack_from("<QQQ", self.b, p + 8)
        return self._at(data), size

    def _name(self, heap_data: int, off: int) -> str:
        end = self.b.index(b"\x00", heap_data + off)
        return self.b[heap_data + off:end].decode("utf-8")

    def _btree_entries(self, addr: int, heap_data: int, out: List[Tuple[str, int]]):
        p = self._at(addr)
        if self.b[p:p + 4] != b"TREE":
            raise H5LiteError("B-tree node signature missing")
        ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
        if ntype != 0:
            raise H5LiteError("group B-tree expected (node type 0)")
        q = p + 24  # after signature, type, level, entries used, left / right sibling
        for i in range(used):
            child = struct.unpack_from("<Q", self.b, q + 8 + i * 16)[0]  # key_i, child_i, key_{i+1}, ...
            if level > 0:
                self._btree_entries(child, heap_data, out)
            else:
                s = self._at(child)
                if self.b[s:s + 4] != b"SNOD":
                    raise H5LiteError("symbol-table node signature missing")
                nsym = struct.unpack_from("<H", self.b, s + 6)[0]
                for k in range(nsym):
                    e = self._sym_entry(s + 8 + 40 * k)
                    out.append((self._name(heap_data, e["name_off"]), e["ohdr"]))

    def _messages(self, ohdr_addr: int):
        p = self._at(ohdr_addr)
        if self.b[p:p + 4] == b"OHDR":
            raise H5LiteError("version-2 object headers are not supported")
        ver, _, nmsg, _refs, size = struct.unpack_from("<BBHII", self.b, p)
        if ver != 1:
            raise H5LiteError(f"object header version {ver} is not supported")
        blocks = [(p + 16, size)]  # the first block starts 8-byte aligned after the 12-byte prefix
        msgs = []
        while blocks and len(msgs) < nmsg:
            q, left = blocks.pop(0)
            while left >= 8 and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.b, q)
                body = self.b[q + 8:q + 8 + msize]
                msgs.append((mtype, body))
                if mtype == 0x0010:  # continuation
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._at(coff), clen))
                q += 8 + msize
                left -= 8 + msize
        return msgs

    # -- objects
    def group_members(self, ohdr_addr: int) -> List[Tuple[str, int]]:
        for mtype, body in self._messages(ohdr_addr):
            if mtype == 0x0011:  # symbol table message: B-tree address, local heap address
                btree, heap = struct.unpack_from("<QQ", body, 0)
                heap_data, _ = self._heap(heap)
                out: List[Tuple[str, int]] = []
                self._btree_entries(btree, heap_data, out)
                return out
            if mtype in (0x0002, 0x0006):
                raise H5LiteError("new-style (link message) groups are not supported")
        raise H5LiteError("object is not a group")

    def is_group(self, ohdr_addr: int) -> bool:
        return any(t == 0x0011 for t, _ in self._messages(ohdr_addr))

    def dataset(self, ohdr_addr: int) -> np.ndarray:
        shape = dtype = None
        layout = None
        for mtype, body in self._messages(ohdr_addr):
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = body[0], body[1], body[2]
                q = 8 if ver == 1 else 4
                if ver not in (1, 2):
                    raise H5LiteError(f"dataspace version {ver}")
                if ver == 2 and body[3] == 2:
                    raise H5LiteError("null dataspace")
                shape = struct.unpack_from("<%dQ" % rank, body, q) if rank else ()
            elif mtype == 0x0003:  # datatype
                cls, ver = body[0] & 15, body[0] >> 4
                bits0 = body[1]
                size = struct.unpack_from("<I", body, 4)[0]
                order = ">" if bits0 & 1 else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
                elif cls == 1:
                    if size not in (2, 4, 8):
                        raise H5LiteError(f"floating-point size {size}")
                    dtype = np.dtype(f"{order}f{size}")
                else:
                    raise H5LiteError(f"datatype class {cls} is not supported")
            elif mtype == 0x0008:  # data layout
                ver = body[0]
                if ver == 3:
                    lclass = body[1]
                    if lclass == 1:
                        addr, size = struct.unpack_from("<QQ", body, 2)
                        layout = ("contiguous", addr, size)
                    elif lclass == 0:
                        size = struct.unpack_from("<H", body, 2)[0]
                        layout = ("compact", bytes(body[4:4 + size]), size)
                    else:
                        raise H5LiteError("chunked datasets are not supported (the reference writes contiguous ones)")
                elif ver in (1, 2):
                    rank, lclass = body[1], body[2]
                    if lclass != 1:
                        raise H5LiteError("only contiguous datasets are supported for layout versions 1 / 2")
                    addr = struct.unpack_from("<Q", body, 8)[0]
                    layout = ("contiguous", addr, None)
                else:
                    raise H5LiteError(f"data layout version {ver}")
        if shape is None or dtype is None or layout is None:
            raise H5LiteError("object is not a (simple) dataset")
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        nbytes = n * dtype.itemsize
        if layout[0] == "compact":
            raw = layout[1][:nbytes]
        else:
            addr = layout[1]
            if addr == UNDEF:  # never allocated: all fill value (zeros)
                return np.zeros(shape, dtype=dtype.newbyteorder("="))
            raw = self.b[self._at(addr):self._at(addr) + nbytes]
        if len(raw) != nbytes:
            raise H5LiteError("truncated dataset")
        arr = np.frombuffer(raw, dtype=dtype).reshape(shape)
        return arr.astype(dtype.newbyteorder("="), copy=True)


def read(path: str) -> Dict[str, np.ndarray]:
    """``{name: array}`` for every dataset of the root group (sub-groups are walked, names joined with '/')."""
    with open(path, "rb") as f:
        rd = _Reader(f.read())
    out: Dict[str, np.ndarray] = {}

    def walk(ohdr, prefix):
        for name, addr in rd.group_members(ohdr):
            if rd.is_group(addr):
                walk(addr, prefix + name + "/")
            else:
                out[prefix + name] = rd.dataset(addr)

    walk(rd.root_entry["ohdr"], "")
    return out


def describe(path: str) -> dict:
    """Structural summary (superblock fields, member names in stored order) — used by the tests to compare the writer's
    layout with a libhdf5-written file."""
    with open(path, "rb") as f:
        rd = _Reader(f.read())
    members = rd.group_members(rd.root_entry["ohdr"])
    return {"base": rd.base, "eof": rd.eof, "leaf_k": rd.leaf_k, "int_k": rd.int_k,
            "root_cache_type": rd.root_entry["cache"], "members": [m[0] for m in members]}


# ======================================================================================================== writer
def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\x00" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _ohdr(msgs: List[bytes]) -> bytes:
    payload = b"".join(msgs)
    # version 1, reserved, number of messages, reference count 1, header size, 4 bytes of padding to the 8-byte boundary
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(payload)) + payload


_DTYPES = {
    np.dtype("<f4"): struct.pack("<B3BI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127),
    np.dtype("<f8"): struct.pack("<B3BI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023),
    np.dtype("<i4"): struct.pack("<B3BI", 0x10, 0x08, 0, 0, 4) + struct.pack("<HH", 0, 32),
    np.dtype("<i8"): struct.pack("<B3BI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64),
    np.dtype("u1"): struct.pack("<B3BI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8),
}


def write(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write ``{name: array}`` as contiguous datasets of a flat root group (names must not contain '/').

    File layout (every address 8-byte aligned, base address 0):
    superblock v0 | root object header (symbol-table message) | local heap + names | B-tree node | one symbol-table
    node holding all entries in name order | per dataset: object header (dataspace, datatype, fill value, layout), raw data.
    The "group leaf node K" of the superblock is raised to ``ceil(n / 2)`` when there are more than 8 datasets, so a single
    symbol-table node (capacity 2K) and a one-entry B-tree hold any number of videos."""
    names = sorted(datasets.keys(), key=lambda s: s.encode("utf-8"))  # libhdf5 orders symbol-table entries by strcmp
    arrays = []
    for nme in names:
        if not nme or "/" in nme or "\x00" in nme:
            raise H5LiteError(f"bad dataset name {nme!r}")
        a = np.ascontiguousarray(datasets[nme])
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        dt = np.dtype(dt.str.replace("=", "<").replace("|", ""))
        if np.dtype(dt) not in _DTYPES:
            raise H5LiteError(f"dtype {a.dtype} is not supported by the minimal writer")
        arrays.append(a.astype(dt, copy=False))
    n = len(names)
    leaf_k = max(4, (n + 1) // 2)
    if leaf_k > 0xFFFF:
        raise H5LiteError("too many datasets for one symbol-table node")
    int_k = 16

    # ---- local heap data segment: offset 0 = "" (the root's own name), then the names, each padded to 8 bytes
    heap_data = bytearray(b"\x00" * 8)
    name_off = []
    for nme in names:
        name_off.append(len(heap_data))
        raw = nme.encode("utf-8") + b"\x00"
        heap_data += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
    # keep one free block at the end, as libhdf5 does (next = 1 "no more blocks", size of the block)
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)

    # ---- addresses
    SB = 96  # superblock v0 with 8-byte offsets: 24 + 4*8 + 40
    root_ohdr_addr = SB
    root_ohdr_len = 16 + (8 + 16) + 8  # symbol-table message + a zero-size NIL message: libhdf5 never writes < 32 bytes of messages
    heap_addr = _pad8(root_ohdr_addr + root_ohdr_len)
    heap_data_addr = heap_addr + 32
    btree_addr = _pad8(heap_data_addr + len(heap_data))
    btree_len = 24 + (2 * int_k + 1) * 8 + 2 * int_k * 8
    snod_addr = btree_addr + btree_len
    snod_len = 8 + 2 * leaf_k * 40
    cur = _pad8(snod_addr + snod_len)

    ds_hdr_addr, ds_data_addr, ds_hdr = [], [], []
    for a in arrays:
        rank = a.ndim
        space = struct.pack("<BBB5x", 1, rank, 0) + struct.pack("<%dQ" % rank, *a.shape)
        # fill value message exactly as libhdf5 writes it for a default dataset: version 1, space allocation "late",
        # fill written "if set", fill value defined with size 0 (= the datatype's default, zero)
        fill = struct.pack("<BBBBI", 1, 2, 2, 1, 0)
        hdr_len = 16 + len(_msg(1, space)) + len(_msg(3, _DTYPES[a.dtype])) + len(_msg(5, fill)) + len(_msg(8, b"\x00" * 18))
        ds_hdr_addr.append(cur)
        data_addr = _pad8(cur + hdr_len)
        ds_data_addr.append(data_addr if a.nbytes else UNDEF)
        layout = struct.pack("<BBQQ", 3, 1, ds_data_addr[-1], a.nbytes)
        ds_hdr.append(_ohdr([_msg(1, space), _msg(3, _DTYPES[a.dtype], flags=1), _msg(5, fill, flags=1), _msg(8, layout)]))
        assert len(ds_hdr[-1]) == hdr_len
        cur = _pad8(data_addr + a.nbytes)
    eof = cur

    out = bytearray(eof)
    # ---- superblock
    root_entry = struct.pack("<QQII", 0, root_ohdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    out[0:SB] = (SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, int_k, 0)
                 + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + root_entry)
    # ---- root group
    hdr = _ohdr([_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr)), _msg(0x0, b"")])
    assert len(hdr) == root_ohdr_len
    out[root_ohdr_addr:root_ohdr_addr + len(hdr)] = hdr
    out[heap_addr:heap_addr + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)
    out[heap_data_addr:heap_data_addr + len(heap_data)] = heap_data
    # B-tree leaf level: keys are heap offsets of names; key[0] = "" (offset 0), key[1] = the largest name of child 0
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if n else 0, UNDEF, UNDEF)
    if n:
        node += struct.pack("<QQQ", 0, snod_addr, name_off[-1])
    out[btree_addr:btree_addr + len(node)] = node
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, n)
    for off, addr in zip(name_off, ds_hdr_addr):
        snod += struct.pack("<QQII16x", off, addr, 0, 0)
    out[snod_addr:snod_addr + len(snod)] = snod
    # ---- datasets
    for a, ha, da, hd in zip(arrays, ds_hdr_addr, ds_data_addr, ds_hdr):
        out[ha:ha + len(hd)] = hd
        if a.nbytes:
            out[da:da + a.nbytes] = a.tobytes()
    with open(path, "wb") as f:
        f.write(out)
