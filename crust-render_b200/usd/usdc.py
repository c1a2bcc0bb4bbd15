"""A reader for binary USD "crate" files (`PXR-USDC`, versions 0.4 - 0.10) producing the same Stage /
Prim / Attribute objects as usda.py, so importer.py handles both.

SURVEY §8(f) rank 1: `samples/PointInstancedMedCity.usd` is a crate file whose arrays are all authored
as single time samples. The reference reads it through the third-party `openusd` crate; this is an
independent restatement of the published crate layout (pxr/usd/usd/crateFile.cpp, integerCoding.cpp):
bootstrap + table of contents; TOKENS / STRINGS / FIELDS / FIELDSETS / PATHS / SPECS sections; LZ4 block
compression behind TfFastCompression's chunk header; Usd_IntegerCompression (delta from the previous value,
2-bit width codes, a common delta); 64-bit ValueReps (type, inlined / array / compressed flags, 48-bit
payload); compressed float arrays ('i' = integers, 't' = lookup table); TimeSamples with their two
forward-offset records. Composition arcs are not composed. Only the value types the samples use are decoded;
anything else comes back as an `Unsupported` marker instead of failing the whole file.
"""
from __future__ import annotations

import struct
from typing import Any, Dict, List, Optional

import numpy as np

from .usda import Attribute, Prim, Stage


class Unsupported:
    def __init__(self, type_id: int):
        self.type_id = type_id

    def __repr__(self):
        return f"<unsupported crate value type {self.type_id}>"


def lz4_block(src: bytes, out_size: int) -> bytes:
    """LZ4 block format."""
    try:
        import pyarrow as pa
        return pa.decompress(src, decompressed_size=out_size, codec="lz4_raw", asbytes=True)
    except Exception:
        pass
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        tok = src[i]
        i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]
        i += lit
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8)
        i += 2
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        if off >= ml:
            out += out[start:start + ml]
        else:
            for k in range(ml):
                out.append(out[start + k])
    return bytes(out)


def fast_decompress(src: bytes, out_size: int) -> bytes:
    """TfFastCompression::DecompressFromBuffer: one byte of chunk count (0 = a single LZ4 block)."""
    n_chunks = src[0]
    if n_chunks == 0:
        return lz4_block(src[1:], out_size)
    out = bytearray()
    p = 1
    for _ in range(n_chunks):
        (sz,) = struct.unpack_from("<i", src, p)
        p += 4
        out += lz4_block(src[p:p + sz], min(out_size - len(out), 127 << 24))
        p += sz
    return bytes(out)


def decode_ints(buf: bytes, count: int, width: int = 4, signed: bool = False) -> np.ndarray:
    """Usd_IntegerCompression after LZ4: common delta, 2-bit codes, variable-width deltas, running sum."""
    if count == 0:
        return np.zeros(0, np.int64)
    if width == 4:
        common = struct.unpack_from("<i", buf, 0)[0]
        sizes, fmts, p = (0, 1, 2, 4), (None, "<b", "<h", "<i"), 4
    else:
        common = struct.unpack_from("<q", buf, 0)[0]
        sizes, fmts, p = (0, 2, 4, 8), (None, "<h", "<i", "<q"), 8
    n_code_bytes = (count * 2 + 7) // 8
    codes = np.frombuffer(buf, np.uint8, n_code_bytes, p)
    p += n_code_bytes
    c = ((codes[:, None] >> np.array([0, 2, 4, 6], np.uint8)) & 3).reshape(-1)[:count]
    deltas = np.empty(count, np.int64)
    deltas[c == 0] = common
    # variable-width values appear in order; gather per code with running offsets
    sz = np.asarray(sizes, np.int64)[c]
    offs = p + np.concatenate([[0], np.cumsum(sz)[:-1]])
    raw = np.frombuffer(buf, np.uint8)
    for code in (1, 2, 3):
        idx = np.nonzero(c == code)[0]
        if len(idx) == 0:
            continue
        w = sizes[code]
        gather = offs[idx][:, None] + np.arange(w)
        vals = raw[gather].copy().view({1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[w]).reshape(-1)
        deltas[idx] = vals
    out = np.cumsum(deltas)
    if width == 4:
        out = out.astype(np.int32) if signed else (out & 0xFFFFFFFF).astype(np.uint32)
    return out.astype(np.int64) if signed or width == 8 else out.astype(np.int64)


# crate type ids (crateDataTypes.h)
T_BOOL, T_UCHAR, T_INT, T_UINT, T_INT64, T_UINT64, T_HALF, T_FLOAT, T_DOUBLE, T_STRING, T_TOKEN, T_ASSET = range(1, 13)
T_MATRIX2D, T_MATRIX3D, T_MATRIX4D, T_QUATD, T_QUATF, T_QUATH = 13, 14, 15, 16, 17, 18
T_VEC2D, T_VEC2F, T_VEC2H, T_VEC2I, T_VEC3D, T_VEC3F, T_VEC3H, T_VEC3I, T_VEC4D, T_VEC4F, T_VEC4H, T_VEC4I = range(19, 31)
T_DICT, T_TOKENLISTOP, T_STRINGLISTOP, T_PATHLISTOP, T_REFLISTOP = 31, 32, 33, 34, 35
T_PATHVECTOR, T_TOKENVECTOR, T_SPECIFIER, T_PERMISSION, T_VARIABILITY, T_VARSELMAP, T_TIMESAMPLES = 40, 41, 42, 43, 44, 45, 46
T_DOUBLEVECTOR, T_STRINGVECTOR, T_VALUEBLOCK = 48, 50, 51

_NUMERIC = {  # type id -> (numpy dtype, components)
    T_UCHAR: ("u1", 1), T_INT: ("<i4", 1), T_UINT: ("<u4", 1), T_INT64: ("<i8", 1), T_UINT64: ("<u8", 1), T_HALF: ("<f2", 1), T_FLOAT: ("<f4", 1),
    T_DOUBLE: ("<f8", 1), T_MATRIX2D: ("<f8", 4), T_MATRIX3D: ("<f8", 9), T_MATRIX4D: ("<f8", 16), T_QUATD: ("<f8", 4), T_QUATF: ("<f4", 4),
    T_QUATH: ("<f2", 4), T_VEC2D: ("<f8", 2), T_VEC2F: ("<f4", 2), T_VEC2H: ("<f2", 2), T_VEC2I: ("<i4", 2), T_VEC3D: ("<f8", 3), T_VEC3F: ("<f4", 3),
    T_VEC3H: ("<f2", 3), T_VEC3I: ("<i4", 3), T_VEC4D: ("<f8", 4), T_VEC4F: ("<f4", 4), T_VEC4H: ("<f2", 4), T_VEC4I: ("<i4", 4),
}
_SPECIFIERS = {0: "def", 1: "over", 2: "class"}


class CrateFile:
    def __init__(self, data: bytes):
        if data[:8] != b"PXR-USDC":
            raise ValueError("not a USD crate file")
        self.d = data
        self.version = tuple(data[8:11])
        if self.version < (0, 4, 0):
            raise ValueError(f"crate version {self.version} predates compressed structural sections; not supported")
        (toc,) = struct.unpack_from("<Q", data, 16)
        (n,) = struct.unpack_from("<Q", data, toc)
        self.sections = {}
        for i in range(n):
            o = toc + 8 + 32 * i
            name = data[o:o + 16].split(b"\0")[0].decode()
            self.sections[name] = struct.unpack_from("<QQ", data, o + 16)
        self._tokens()
        self._strings()
        self._fields()
        self._fieldsets()
        self._paths()
        self._specs()

    # ---- sections ---------------------------------------------------------------------------------
    def _compressed_ints(self, p: int, count: int, width=4, signed=False):
        (csize,) = struct.unpack_from("<Q", self.d, p)
        p += 8
        raw_size = (4 if width == 4 else 8) + (count * 2 + 7) // 8 + count * width
        buf = fast_decompress(self.d[p:p + csize], raw_size)
        return decode_ints(buf, count, width, signed), p + csize

    def _tokens(self):
        p, _ = self.sections["TOKENS"]
        n, usize, csize = struct.unpack_from("<QQQ", self.d, p)
        blob = fast_decompress(self.d[p + 24:p + 24 + csize], usize)
        self.tokens = [t.decode("utf-8", "replace") for t in blob.split(b"\0")[:n]]

    def _strings(self):
        p, _ = self.sections["STRINGS"]
        (n,) = struct.unpack_from("<Q", self.d, p)
        self.strings = [self.tokens[i] for i in struct.unpack_from(f"<{n}I", self.d, p + 8)]

    def _fields(self):
        p, _ = self.sections["FIELDS"]
        (n,) = struct.unpack_from("<Q", self.d, p)
        tok, p = self._compressed_ints(p + 8, n)
        (csize,) = struct.unpack_from("<Q", self.d, p)
        reps = np.frombuffer(fast_decompress(self.d[p + 8:p + 8 + csize], n * 8), "<u8")
        self.fields = [(self.tokens[int(t)], int(r)) for t, r in zip(tok, reps)]

    def _fieldsets(self):
        p, _ = self.sections["FIELDSETS"]
        (n,) = struct.unpack_from("<Q", self.d, p)
        self.fieldsets, _ = self._compressed_ints(p + 8, n)

    def _paths(self):
        p, _ = self.sections["PATHS"]
        (n_paths,) = struct.unpack_from("<Q", self.d, p)
        (n_enc,) = struct.unpack_from("<Q", self.d, p + 8)
        idx, q = self._compressed_ints(p + 16, n_enc)
        elem, q = self._compressed_ints(q, n_enc, signed=True)
        jumps, q = self._compressed_ints(q, n_enc, signed=True)
        self.paths: List[Optional[str]] = [None] * n_paths
        stack = [(0, "")]
        while stack:
            cur, parent = stack.pop()
            while True:
                this = cur
                cur += 1
                if parent == "":
                    path = "/"
                else:
                    e = int(elem[this])
                    tok = self.tokens[abs(e)]
                    path = parent + "." + tok if e < 0 else (parent.rstrip("/") + "/" + tok)
                self.paths[int(idx[this])] = path
                j = int(jumps[this])
                has_child, has_sibling = (j > 0 or j == -1), j >= 0
                if has_child:
                    if has_sibling:
                        stack.append((this + j, parent))
                    parent = path
                    continue
                if has_sibling:
                    continue
                break

    def _specs(self):
        p, _ = self.sections["SPECS"]
        (n,) = struct.unpack_from("<Q", self.d, p)
        pi, q = self._compressed_ints(p + 8, n)
        fs, q = self._compressed_ints(q, n)
        st, q = self._compressed_ints(q, n)
        self.specs = [(self.paths[int(a)], int(b), int(c)) for a, b, c in zip(pi, fs, st)]

    # ---- values -----------------------------------------------------------------------------------
    def spec_fields(self, fieldset_index: int) -> Dict[str, int]:
        out = {}
        i = fieldset_index
        while self.fieldsets[i] != 0xFFFFFFFF:
            name, rep = self.fields[int(self.fieldsets[i])]
            out[name] = rep
            i += 1
        return out

    def value(self, rep: int) -> Any:
        t = (rep >> 48) & 0xFF
        is_array, inlined, compressed = bool(rep >> 63 & 1), bool(rep >> 62 & 1), bool(rep >> 61 & 1)
        payload = rep & ((1 << 48) - 1)
        d = self.d
        if inlined:
            if t == T_BOOL:
                return bool(payload & 1)
            if t in (T_UCHAR, T_UINT):
                return payload & 0xFFFFFFFF
            if t == T_INT:
                return struct.unpack("<i", struct.pack("<I", payload & 0xFFFFFFFF))[0]
            if t == T_FLOAT:
                return struct.unpack("<f", struct.pack("<I", payload & 0xFFFFFFFF))[0]
            if t == T_DOUBLE:  # stored as a float when exactly representable
                return float(struct.unpack("<f", struct.pack("<I", payload & 0xFFFFFFFF))[0])
            if t == T_HALF:
                return float(np.frombuffer(struct.pack("<H", payload & 0xFFFF), "<f2")[0])
            if t in (T_INT64, T_UINT64):
                return struct.unpack("<i", struct.pack("<I", payload & 0xFFFFFFFF))[0]
            if t == T_TOKEN:
                return self.tokens[payload]
            if t == T_STRING:
                return self.strings[payload]
            if t == T_ASSET:
                return self.tokens[payload]
            if t == T_SPECIFIER:
                return _SPECIFIERS.get(payload, "def")
            if t == T_VARIABILITY:
                return "uniform" if payload == 1 else "varying"
            if t in (T_VEC2I, T_VEC3I, T_VEC4I, T_VEC2F, T_VEC3F, T_VEC4F, T_VEC2D, T_VEC3D, T_VEC4D, T_VEC2H, T_VEC3H, T_VEC4H):
                n = _NUMERIC[t][1]
                b = struct.pack("<Q", payload)[:n]
                return tuple(float(x) for x in np.frombuffer(b, np.int8))  # small integral components
            if t == T_MATRIX4D:  # diagonal stored as 4 int8
                m = np.zeros((4, 4))
                m[np.arange(4), np.arange(4)] = np.frombuffer(struct.pack("<Q", payload)[:4], np.int8)
                return m
            if t == T_QUATH or t == T_QUATF or t == T_QUATD:
                return Unsupported(t)
            if t == T_DICT and payload == 0:
                return {}
            if t == T_VALUEBLOCK:
                return None
            if is_array and payload == 0:
                return np.zeros((0,), _NUMERIC.get(t, ("<f4", 1))[0])
            return Unsupported(t)
        if is_array:
            if payload == 0:  # an empty array has no data block
                return [] if t == T_TOKEN else np.zeros((0, _NUMERIC[t][1]) if t in _NUMERIC and _NUMERIC[t][1] > 1 else (0,), _NUMERIC.get(t, ("<f4", 1))[0])
            return self._array(t, payload, compressed)
        p = payload
        if t in _NUMERIC:
            dt, n = _NUMERIC[t]
            a = np.frombuffer(d, dt, n, p)
            if t == T_MATRIX4D:
                return a.reshape(4, 4).astype(np.float64)
            if t in (T_QUATD, T_QUATF, T_QUATH):
                return (float(a[3]), float(a[0]), float(a[1]), float(a[2]))
            return tuple(float(x) for x in a) if n > 1 else a[0].item()
        if t == T_TOKENVECTOR:
            (n,) = struct.unpack_from("<Q", d, p)
            return [self.tokens[i] for i in struct.unpack_from(f"<{n}I", d, p + 8)]
        if t == T_PATHVECTOR:
            (n,) = struct.unpack_from("<Q", d, p)
            return [self.paths[i] for i in struct.unpack_from(f"<{n}I", d, p + 8)]
        if t == T_DOUBLEVECTOR:
            (n,) = struct.unpack_from("<Q", d, p)
            return list(struct.unpack_from(f"<{n}d", d, p + 8))
        if t == T_PATHLISTOP:
            return self._path_list_op(p)
        if t == T_TOKENLISTOP:
            return self._token_list_op(p)
        if t == T_TIMESAMPLES:
            return self._time_samples(p)
        if t == T_DICT:
            try:
                return self._dictionary(p)
            except Exception:  # metadata only (assetInfo, customData): never needed by the importer
                return Unsupported(t)
        return Unsupported(t)

    def _array(self, t: int, p: int, compressed: bool):
        d = self.d
        if self.version >= (0, 7, 0):
            (n,) = struct.unpack_from("<Q", d, p)
            p += 8
        else:
            (n,) = struct.unpack_from("<I", d, p)
            p += 4
        if t == T_TOKEN:
            return [self.tokens[i] for i in struct.unpack_from(f"<{n}I", d, p)]
        if t not in _NUMERIC:
            return Unsupported(t)
        dt, comps = _NUMERIC[t]
        if compressed and t in (T_INT, T_UINT, T_INT64, T_UINT64):
            w = 8 if t in (T_INT64, T_UINT64) else 4
            vals, _ = self._compressed_ints(p, n, w, signed=t in (T_INT, T_INT64))
            return vals.astype(dt)
        if compressed and t in (T_HALF, T_FLOAT, T_DOUBLE):
            code = d[p:p + 1]
            p += 1
            if code == b"i":
                vals, _ = self._compressed_ints(p, n, 4, signed=True)
                return vals.astype(dt)
            if code == b"t":
                (lut_n,) = struct.unpack_from("<I", d, p)
                lut = np.frombuffer(d, dt, lut_n, p + 4)
                idx, _ = self._compressed_ints(p + 4 + lut_n * np.dtype(dt).itemsize, n)
                return lut[idx]
            raise ValueError("corrupt compressed float array")
        a = np.frombuffer(d, dt, n * comps, p)
        if t in (T_QUATD, T_QUATF, T_QUATH):  # stored (i, j, k, real); hand out (w, x, y, z) like the text format
            return a.reshape(n, 4)[:, [3, 0, 1, 2]]
        return a.reshape(n, comps) if comps > 1 else a

    def _list_op(self, p: int, read_items):
        hdr = self.d[p]
        p += 1
        out = {}
        for bit, key in ((1, "explicit"), (2, "added"), (4, "deleted"), (8, "ordered"), (16, "prepended"), (32, "appended")):
            if key == "explicit":
                if hdr & 1:
                    out["is_explicit"] = True
                continue
        # items follow in the order: explicit, added, prepended, appended, deleted, ordered
        for flag, key in ((2, "explicit"), (4, "added"), (32, "prepended"), (64, "appended"), (8, "deleted"), (16, "ordered")):
            if hdr & flag:
                items, p = read_items(p)
                out[key] = items
        return out

    def _path_list_op(self, p: int):
        def items(q):
            (n,) = struct.unpack_from("<Q", self.d, q)
            return [self.paths[i] for i in struct.unpack_from(f"<{n}I", self.d, q + 8)], q + 8 + 4 * n
        return self._list_op(p, items)

    def _token_list_op(self, p: int):
        def items(q):
            (n,) = struct.unpack_from("<Q", self.d, q)
            return [self.tokens[i] for i in struct.unpack_from(f"<{n}I", self.d, q + 8)], q + 8 + 4 * n
        return self._list_op(p, items)

    def _dictionary(self, p: int):
        (n,) = struct.unpack_from("<Q", self.d, p)
        p += 8
        out = {}
        for _ in range(n):
            (si,) = struct.unpack_from("<I", self.d, p)
            p += 4
            (off,) = struct.unpack_from("<q", self.d, p)
            (rep,) = struct.unpack_from("<Q", self.d, p + off)
            out[self.strings[si]] = self.value(rep)
            p += 8
        return out

    def _time_samples(self, p: int) -> Dict[float, Any]:
        d = self.d
        (off,) = struct.unpack_from("<q", d, p)
        q = p + off
        (times_rep,) = struct.unpack_from("<Q", d, q)
        times = self.value(times_rep)
        if not isinstance(times, (list, np.ndarray)):
            times = [times]
        q += 8
        (off2,) = struct.unpack_from("<q", d, q)
        q += off2
        (n,) = struct.unpack_from("<Q", d, q)
        reps = struct.unpack_from(f"<{n}Q", d, q + 8)
        return {float(t): self.value(r) for t, r in zip(times, reps)}


SPEC_ATTRIBUTE, SPEC_PRIM, SPEC_PSEUDOROOT, SPEC_RELATIONSHIP = 1, 6, 7, 8


def parse(data: bytes) -> Stage:
    cf = CrateFile(data)
    by_path = {path: (cf.spec_fields(fs), st) for path, fs, st in cf.specs}
    prims: Dict[str, Prim] = {}
    meta: Dict[str, Any] = {}

    def make_prim(path: str, parent: Optional[Prim]) -> Prim:
        fields, _ = by_path[path]
        name = path.rsplit("/", 1)[1]
        p = Prim(cf.value(fields["specifier"]) if "specifier" in fields else "over", "", name, path, parent=parent)
        if "typeName" in fields:
            p.type = cf.value(fields["typeName"])
        for k in ("active", "instanceable", "kind", "apiSchemas"):
            if k in fields:
                p.metadata[k] = cf.value(fields[k])
        for prop in (cf.value(fields["properties"]) if "properties" in fields else []):
            pf, st = by_path.get(path + "." + prop, ({}, 0))
            if st == SPEC_RELATIONSHIP:
                targets = []
                if "targetPaths" in pf:
                    lo = cf.value(pf["targetPaths"])
                    for key in ("explicit", "prepended", "appended", "added"):
                        targets += lo.get(key, []) if isinstance(lo, dict) else []
                p.relationships[prop] = targets
            elif st == SPEC_ATTRIBUTE:
                a = Attribute(prop, cf.value(pf["typeName"]) if "typeName" in pf else "", custom=bool(cf.value(pf["custom"])) if "custom" in pf else False)
                if "default" in pf:
                    a.value = cf.value(pf["default"])
                if "timeSamples" in pf:
                    a.time_samples = cf.value(pf["timeSamples"])
                if "connectionPaths" in pf:
                    lo = cf.value(pf["connectionPaths"])
                    if isinstance(lo, dict):
                        tg = [t for key in ("explicit", "prepended", "appended", "added") for t in lo.get(key, [])]
                        if tg:
                            a.metadata["connect"] = tg[0]
                p.attributes[prop] = a
        for child in (cf.value(fields["primChildren"]) if "primChildren" in fields else []):
            cp = (path if path != "/" else "") + "/" + child
            if cp in by_path:
                p.children.append(make_prim(cp, p))
        prims[path] = p
        return p

    root_fields, _ = by_path["/"]
    for k, rep in root_fields.items():
        if k != "primChildren":
            meta[k] = cf.value(rep)
    roots = [make_prim("/" + c, None) for c in (cf.value(root_fields["primChildren"]) if "primChildren" in root_fields else []) if "/" + c in by_path]
    return Stage(meta, roots)


def load(path: str) -> Stage:
    with open(path, "rb") as f:
        return parse(f.read())
