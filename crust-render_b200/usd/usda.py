"""A small reader for the text USD (`.usda`) subset crust-render's bundled samples use.

SURVEY §8(f) rank 1 (scene ingest). The reference reads stages through the third-party `openusd`
crate (crates/crust-core/Cargo.toml:22); this is an independent reader of the crate-free text format,
enough for prim hierarchies, typed attributes (scalars, tuples, arrays, matrices, tokens, asset
paths), relationships, `custom` attributes, prim metadata (`active`, `instanceable`, `kind`,
`apiSchemas`) and `.timeSamples` (the earliest sample stands in for the default, like
`attr.get()` without a time code does for single-sample attributes). Composition arcs
(references, payloads, variants, sublayers) are parsed here — external targets as `Ref(asset, path)`, variant sets as
`Prim.variant_sets` — and composed by `usd/compose.py`.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np


@dataclass
class Attribute:
    name: str
    type: str
    value: Any = None
    custom: bool = False
    uniform: bool = False
    time_samples: Optional[Dict[float, Any]] = None
    metadata: Dict[str, Any] = field(default_factory=dict)

    def get(self):
        """Value at the default time; falls back to the earliest time sample when no default is authored — or
        when the authored default is an empty array while samples exist (PointInstancedMedCity.usd authors
        `positions = []` next to a 40 000-element time sample; what `openusd` returns there is unpinned)."""
        v = self.value
        empty = v is not None and hasattr(v, "__len__") and not isinstance(v, (str, dict)) and len(v) == 0
        if v is not None and not (empty and self.time_samples):
            return v
        if self.time_samples:
            return self.time_samples[min(self.time_samples)]
        return v


@dataclass(frozen=True)
class Ref:
    """A reference / payload target: `@layer.usda@</Prim>` (either part may be empty: an internal path, or a layer's
    defaultPrim)."""
    asset: str
    path: str = ""


@dataclass
class Prim:
    specifier: str          # def | over | class
    type: str               # "" for untyped
    name: str
    path: str
    metadata: Dict[str, Any] = field(default_factory=dict)
    attributes: Dict[str, Attribute] = field(default_factory=dict)
    relationships: Dict[str, List[str]] = field(default_factory=dict)
    children: List["Prim"] = field(default_factory=list)
    parent: Optional["Prim"] = None
    # variantSet name -> variant name -> the variant's body as an `over` of this prim (properties, children, metadata)
    variant_sets: Dict[str, Dict[str, "Prim"]] = field(default_factory=dict)

    def attr(self, name: str):
        a = self.attributes.get(name)
        return None if a is None else a.get()

    def is_active(self) -> bool:
        v = self.metadata.get("active", True)
        return v.strip().lower() != "false" if isinstance(v, str) else bool(v)  # the text reader keeps `false` as a bare word

    def is_abstract(self) -> bool:
        return self.specifier == "class"

    def walk(self):
        yield self
        for c in self.children:
            yield from c.walk()


@dataclass
class Stage:
    metadata: Dict[str, Any]
    roots: List[Prim]

    def prim_at(self, path: str) -> Optional[Prim]:
        parts = [p for p in path.strip("<>").split("/") if p]
        cur = None
        level = self.roots
        for p in parts:
            cur = next((c for c in level if c.name == p), None)
            if cur is None:
                return None
            level = cur.children
        return cur

    def walk(self):
        for r in self.roots:
            yield from r.walk()


# ---- a parsed stage as JSON (tests/golden/*.stage.json: what the GPU box imports, where the sample files are absent) ----
def _enc(v):
    if isinstance(v, np.ndarray):
        return {"__nd__": str(v.dtype), "shape": list(v.shape), "data": v.ravel().tolist()}
    if isinstance(v, np.generic):
        return v.item()
    if isinstance(v, tuple):
        return {"__t__": [_enc(x) for x in v]}
    if isinstance(v, list):
        return [_enc(x) for x in v]
    if isinstance(v, dict):
        return {"__d__": [[_enc(k), _enc(x)] for k, x in v.items()]}
    return v


def _dec(v):
    if isinstance(v, dict):
        if "__nd__" in v:
            return np.asarray(v["data"], dtype=v["__nd__"]).reshape(v["shape"])
        if "__t__" in v:
            return tuple(_dec(x) for x in v["__t__"])
        return {_dec(k): _dec(x) for k, x in v["__d__"]}
    if isinstance(v, list):
        return [_dec(x) for x in v]
    return v


def stage_to_json(stage: Stage) -> dict:
    def prim(p: Prim):
        return {"specifier": p.specifier, "type": p.type, "name": p.name, "metadata": _enc(p.metadata),
                "attributes": [[a.name, a.type, _enc(a.value), a.custom, a.uniform, _enc(a.time_samples), _enc(a.metadata)] for a in p.attributes.values()],
                "relationships": _enc(p.relationships), "children": [prim(c) for c in p.children]}
    return {"metadata": _enc(stage.metadata), "roots": [prim(r) for r in stage.roots]}


def stage_from_json(doc: dict) -> Stage:
    def prim(d, parent):
        p = Prim(d["specifier"], d["type"], d["name"], (parent.path if parent else "") + "/" + d["name"], metadata=_dec(d["metadata"]), parent=parent)
        for name, ty, value, custom, uniform, ts, meta in d["attributes"]:
            p.attributes[name] = Attribute(name, ty, _dec(value), custom, uniform, _dec(ts), _dec(meta))
        p.relationships = _dec(d["relationships"])
        p.children = [prim(c, p) for c in d["children"]]
        return p
    return Stage(_dec(doc["metadata"]), [prim(r, None) for r in doc["roots"]])


_TOKEN = re.compile(r'''
    (?P<ws>\s+|\#[^\n]*)                         |
    (?P<str3>"""(?:.|\n)*?"""|\'\'\'(?:.|\n)*?\'\'\') |
    (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')   |
    (?P<asset>@@@(?:.|\n)*?@@@|@[^@\n]*@)         |
    (?P<path><[^>\n]*>)                           |
    (?P<num>[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)(?![A-Za-z_]))  |
    (?P<ident>[A-Za-z_][A-Za-z0-9_:.\-!]*(?:\[\])?) |
    (?P<punct>[(){}\[\]=,;])
''', re.X)


class _Lexer:
    def __init__(self, text: str):
        self.text = text
        self.pos = 0
        self.peeked = None

    def _next(self):
        n = len(self.text)
        while self.pos < n:
            m = _TOKEN.match(self.text, self.pos)
            if not m:
                raise SyntaxError(f"usda: unexpected character {self.text[self.pos]!r} at offset {self.pos}")
            self.pos = m.end()
            kind = m.lastgroup
            if kind == "ws":
                continue
            return kind, m.group(kind)
        return "eof", ""

    def peek(self):
        if self.peeked is None:
            self.peeked = self._next()
        return self.peeked

    def next(self):
        t = self.peek()
        self.peeked = None
        return t

    def expect(self, value: str):
        k, v = self.next()
        if v != value:
            raise SyntaxError(f"usda: expected {value!r}, got {v!r} near offset {self.pos}")

    def accept(self, value: str) -> bool:
        if self.peek()[1] == value and self.peek()[0] in ("punct", "ident"):
            self.next()
            return True
        return False


_NUMERIC_ARRAY = re.compile(r"[\s\d.eE+\-,()infa]*\]")


def _unquote(s: str) -> str:
    if s.startswith('"""') or s.startswith("'''"):
        return s[3:-3]
    return bytes(s[1:-1], "utf-8").decode("unicode_escape") if "\\" in s else s[1:-1]


class _Parser:
    def __init__(self, text: str):
        if not text.lstrip().startswith("#usda"):
            raise SyntaxError("not a .usda file (missing #usda header)")
        self.lx = _Lexer(text)

    # ---- values ----------------------------------------------------------------------------
    def value(self):
        k, v = self.lx.peek()
        if v == "[":
            return self.array()
        if v == "(":
            self.lx.next()
            items = []
            while not self.lx.accept(")"):
                items.append(self.value())
                self.lx.accept(",")
            return tuple(items)
        if v == "{":
            return self.dictionary()
        self.lx.next()
        if k == "num":
            return float(v) if any(c in v for c in ".eEinfa") else int(v)
        if k in ("str", "str3"):
            return _unquote(v)
        if k == "asset":
            a = v.strip("@")
            if self.lx.peek()[0] == "path":  # @layer@</Prim>: an external composition target
                tgt = Ref(a, self.lx.next()[1].strip("<>"))
                if self.lx.peek()[1] == "(":
                    self.metadata_block()  # layer offset: not applied (no animation is read)
                return tgt
            return a
        if k == "path":
            return v
        if k == "ident":
            if v == "None":
                return None
            return v
        raise SyntaxError(f"usda: unexpected token {v!r} in a value")

    def array(self):
        # fast path for big numeric arrays: hand the bracketed text to numpy
        lx = self.lx
        start = lx.pos - 0
        assert lx.peek()[1] == "["
        # position of the '[' just peeked
        text = lx.text
        open_pos = text.rfind("[", 0, lx.pos)
        m = _NUMERIC_ARRAY.match(text, open_pos + 1)
        if m and m.end() - open_pos > 256:
            body = text[open_pos + 1:m.end() - 1]
            width = 1
            first = body.find("(")
            if first >= 0:
                close = body.find(")", first)
                width = body[first:close].count(",") + 1
            flat = np.array(body.replace("(", " ").replace(")", " ").replace(",", " ").split(), dtype=np.float64)
            lx.peeked = None
            lx.pos = m.end()
            return flat.reshape(-1, width) if width > 1 else flat
        lx.next()
        items = []
        while not lx.accept("]"):
            items.append(self.value())
            lx.accept(",")
        return items

    def dictionary(self):
        lx = self.lx
        lx.expect("{")
        out = {}
        while not lx.accept("}"):
            k, v = lx.next()
            if k in ("str", "str3"):           # timeSamples-style  "key": value   is handled by the caller
                key = _unquote(v)
            elif k == "num":                   # time sample:  1.0: value,
                key = float(v)
                # ':' is not a token of its own; times are written `0: value` -> ident lexing never sees it
            else:
                # typed entry:  <type> <name> = <value>
                if v == "dictionary":
                    name = lx.next()[1]
                    lx.expect("=")
                    out[name] = self.dictionary()
                    continue
                name_k, name = lx.next()
                key = _unquote(name) if name_k in ("str", "str3") else name
            lx.expect("=")
            out[key] = self.value()
            lx.accept(",")
            lx.accept(";")
        return out

    def time_samples(self):
        """{ 0: value, 1: value, }  — ':' is not a token, so parse from the raw text."""
        lx = self.lx
        lx.expect("{")
        out = {}
        while True:
            # skip whitespace / comments manually
            m = re.compile(r"(?:\s+|#[^\n]*)*").match(lx.text, lx.pos)
            lx.pos = m.end()
            if lx.text[lx.pos] == "}":
                lx.pos += 1
                break
            m = re.compile(r"([-+]?[\d.eE+\-]+)\s*:").match(lx.text, lx.pos)
            if not m:
                raise SyntaxError(f"usda: bad timeSamples entry at offset {lx.pos}")
            lx.pos = m.end()
            out[float(m.group(1))] = self.value()
            lx.accept(",")
        return out

    def metadata_block(self):
        lx = self.lx
        out = {}
        lx.expect("(")
        while not lx.accept(")"):
            k, v = lx.next()
            if k in ("str", "str3"):  # bare doc string
                out["comment"] = _unquote(v)
                continue
            key = v
            if key in ("prepend", "append", "delete", "add", "reorder"):
                key = lx.next()[1]
            elif lx.peek()[0] == "ident" and lx.peek()[1] not in ("=",):
                # typed metadata:  <type> <name> = value
                nxt = lx.peek()[1]
                if nxt != "=":
                    if key == "dictionary":
                        name = lx.next()[1]
                        lx.expect("=")
                        out[name] = self.dictionary()
                        continue
                    key = lx.next()[1]
            lx.expect("=")
            out[key] = self.value()
            lx.accept(";")
        return out

    # ---- prims -------------------------------------------------------------------------------
    def prim(self, parent: Optional[Prim]) -> Prim:
        lx = self.lx
        spec = lx.next()[1]
        k, v = lx.next()
        ptype = ""
        if k == "ident":
            ptype = v
            k, v = lx.next()
        name = _unquote(v)
        path = (parent.path if parent else "") + "/" + name
        p = Prim(spec, ptype, name, path, parent=parent)
        if lx.peek()[1] == "(":
            p.metadata = self.metadata_block()
        lx.expect("{")
        while not lx.accept("}"):
            k, v = lx.peek()
            if v in ("def", "over", "class"):
                p.children.append(self.prim(p))
            elif v == "variantSet":
                self.variant_set(p)
            else:
                self.property(p)
        return p

    def variant_set(self, p: Prim):
        """variantSet "name" = { "variant" (metadata) { properties and child prims } ... }"""
        lx = self.lx
        lx.next()
        set_name = _unquote(lx.next()[1])
        lx.expect("=")
        lx.expect("{")
        variants = p.variant_sets.setdefault(set_name, {})
        while not lx.accept("}"):
            k, v = lx.next()
            if k == "eof":
                raise SyntaxError("usda: unterminated variantSet")
            body = Prim("over", "", p.name, p.path, parent=p.parent)
            if lx.peek()[1] == "(":
                body.metadata = self.metadata_block()
            lx.expect("{")
            while not lx.accept("}"):
                k2, v2 = lx.peek()
                if v2 in ("def", "over", "class"):
                    child = self.prim(body)
                    body.children.append(child)
                elif v2 == "variantSet":
                    self.variant_set(body)
                else:
                    self.property(body)
            variants[_unquote(v)] = body

    def property(self, p: Prim):
        lx = self.lx
        custom = uniform = False
        k, v = lx.next()
        while v in ("custom", "uniform", "varying", "prepend", "append", "delete", "add"):
            custom |= v == "custom"
            uniform |= v == "uniform"
            k, v = lx.next()
        if v == "rel":
            name = lx.next()[1]
            targets: List[str] = []
            if lx.accept("="):
                val = self.value()
                targets = [val] if isinstance(val, str) else [t for t in (val or []) if isinstance(t, str)]
            if lx.peek()[1] == "(":
                self.metadata_block()
            p.relationships[name] = [t.strip("<>") for t in targets]
            return
        atype = v
        name = lx.next()[1]
        is_ts = name.endswith(".timeSamples")
        is_connect = name.endswith(".connect")
        base = name.rsplit(".", 1)[0] if (is_ts or is_connect) else name
        a = p.attributes.get(base) or Attribute(base, atype, custom=custom, uniform=uniform)
        if lx.accept("="):
            if is_ts:
                a.time_samples = self.time_samples()
            elif is_connect:
                val = self.value()
                a.metadata["connect"] = val.strip("<>") if isinstance(val, str) else val
            else:
                a.value = self.value()
        if lx.peek()[1] == "(":
            a.metadata.update(self.metadata_block())
        p.attributes[base] = a

    def stage(self) -> Stage:
        lx = self.lx
        meta = self.metadata_block() if lx.peek()[1] == "(" else {}
        roots = []
        while lx.peek()[0] != "eof":
            roots.append(self.prim(None))
        return Stage(meta, roots)


def parse(text: str) -> Stage:
    return _Parser(text).stage()


def load(path: str) -> Stage:
    with open(path, "r", encoding="utf-8") as f:
        return parse(f.read())
