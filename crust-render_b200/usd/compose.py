"""Composition of a USD stage from its layers, as far as crust-render's importer needs it.

The reference opens stages through `openusd` with every payload loaded (`Stage::builder().load(InitialLoadSet::LoadAll)`,
crates/crust-core/src/scene/usd_import.rs:271-300) and walks the COMPOSED prims; `usda.py` / `usdc.py` read one layer each.
This module flattens the arcs that pull other layers in, strongest opinion first (USD's LIVRPS order, of which inherits and
specializes stay with the importer — it resolves them for native instancing, usd_import.rs:166-188):

  * local opinions of the layer stack: the root layer over its `subLayers`, depth first, in listed order;
  * the selected variant of every `variantSet` (`variants = { string set = "choice" }`; without a selection the set
    contributes nothing, like USD without a fallback);
  * `references`, then `payload`, to prims of other layer files (`@file.usda@</Prim>`, or the file's `defaultPrim`), each
    target composed recursively before it is grafted; list-op forms (`prepend references = [...]`) are honoured in listed order.
    Internal references (`</Prim>` without a file) are left in place for the importer.

Grafting a target under a prim copies its properties, metadata and children where the stronger side has no opinion and
re-roots every path inside the copy (relationship targets, `.connect` sources, internal arcs) from the target's namespace to
the destination's. `.usda` and `.usdc` layers mix freely. Cycles of references are an error.
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Optional, Tuple

from . import usda
from .usda import Attribute, Prim, Ref, Stage

_ARC_KEYS = ("references", "payload")


def load_layer(path: str) -> Stage:
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic == b"PXR-USDC":
        from . import usdc
        return usdc.load(path)
    return usda.load(path)


# ---- copying and re-rooting ------------------------------------------------------------------------------------------
def _remap_path(p, src: str, dst: str):
    if isinstance(p, str):
        q = p.strip("<>")
        if q == src or q.startswith(src + "/") or q.startswith(src + "."):
            out = dst + q[len(src):]
            return "<" + out + ">" if p.startswith("<") else out
    return p


def _remap_value(v, src, dst):
    if isinstance(v, str):
        return _remap_path(v, src, dst) if v.lstrip("<").startswith("/") else v
    if isinstance(v, list):
        return [_remap_value(x, src, dst) for x in v]
    if isinstance(v, tuple):
        return tuple(_remap_value(x, src, dst) for x in v)
    if isinstance(v, Ref) and not v.asset:
        return Ref("", _remap_path(v.path, src, dst))
    return v


def _clone(prim: Prim, parent: Optional[Prim], name: str, src_root: str, dst_root: str) -> Prim:
    """Deep copy of `prim` as a child of `parent` called `name`, with every path under src_root re-rooted to dst_root
    (variant bodies included: they are resolved after the copy, in the destination's namespace)."""
    path = (parent.path if parent else "") + "/" + name
    out = Prim(prim.specifier, prim.type, name, path, parent=parent)
    out.metadata = {k: _remap_value(copy.deepcopy(v), src_root, dst_root) for k, v in prim.metadata.items()}
    for k, a in prim.attributes.items():
        b = Attribute(a.name, a.type, a.value, a.custom, a.uniform, a.time_samples, dict(a.metadata))
        if "connect" in b.metadata:
            b.metadata["connect"] = _remap_value(b.metadata["connect"], src_root, dst_root)
        out.attributes[k] = b
    out.relationships = {k: [_remap_path(t, src_root, dst_root) for t in v] for k, v in prim.relationships.items()}
    out.variant_sets = {sn: {vn: _clone(body, parent, name, src_root, dst_root) for vn, body in vs.items()} for sn, vs in prim.variant_sets.items()}
    out.children = [_clone(c, out, c.name, src_root, dst_root) for c in prim.children]
    return out


def _merge(strong: Prim, weak: Prim, src_root: str, dst_root: str, arcs_too: bool = False) -> None:
    """Folds `weak`'s opinions into `strong` wherever `strong` has none. `weak` lives at src_root, `strong` at dst_root."""
    if not strong.type:
        strong.type = weak.type
    if strong.specifier == "over" and weak.specifier != "over":
        strong.specifier = weak.specifier
    for k, v in weak.metadata.items():
        if k in _ARC_KEYS and not arcs_too:
            continue
        if k == "variants" and isinstance(v, dict) and isinstance(strong.metadata.get(k), dict):
            for sn, choice in v.items():  # selections resolve per set: the stronger side's choice of one set does not hide the others
                strong.metadata[k].setdefault(sn, choice)
        elif k not in strong.metadata:
            strong.metadata[k] = _remap_value(copy.deepcopy(v), src_root, dst_root)
    for k, a in weak.attributes.items():
        s = strong.attributes.get(k)
        if s is None:
            b = Attribute(a.name, a.type, a.value, a.custom, a.uniform, a.time_samples, dict(a.metadata))
            if "connect" in b.metadata:
                b.metadata["connect"] = _remap_value(b.metadata["connect"], src_root, dst_root)
            strong.attributes[k] = b
        else:  # value-level resolution: the stronger layer may author only the metadata, or only the time samples
            if s.value is None and s.time_samples is None:
                s.value, s.time_samples = a.value, a.time_samples
            for mk, mv in a.metadata.items():
                s.metadata.setdefault(mk, _remap_value(mv, src_root, dst_root) if mk == "connect" else mv)
    for k, v in weak.relationships.items():
        if k not in strong.relationships:
            strong.relationships[k] = [_remap_path(t, src_root, dst_root) for t in v]
    for set_name, variants in weak.variant_sets.items():
        mine = strong.variant_sets.setdefault(set_name, {})
        for vn, body in variants.items():
            if vn not in mine:
                mine[vn] = _clone(body, strong.parent, strong.name, src_root, dst_root)
    by_name = {c.name: c for c in strong.children}
    for c in weak.children:
        s = by_name.get(c.name)
        if s is None:
            strong.children.append(_clone(c, strong, c.name, src_root, dst_root))
        else:
            _merge(s, c, src_root, dst_root, arcs_too=True)


# ---- the composer ------------------------------------------------------------------------------------------------------
def _anchor(prim: Prim, base: str) -> None:
    """Makes the asset paths of every arc under `prim` absolute (anchored to the directory of the layer that authored them),
    so that a prim grafted into another layer still finds its own targets."""
    for arc in _ARC_KEYS:
        v = prim.metadata.get(arc)
        if v is None:
            continue
        out = []
        for tgt in _as_list(v):
            if isinstance(tgt, Ref) and tgt.asset:
                tgt = Ref(os.path.abspath(os.path.join(base, tgt.asset)), tgt.path)
            elif isinstance(tgt, str) and tgt and not tgt.lstrip("<").startswith("/"):
                tgt = Ref(os.path.abspath(os.path.join(base, tgt)), "")
            out.append(tgt)
        prim.metadata[arc] = out
    for vs in prim.variant_sets.values():
        for body in vs.values():
            _anchor(body, base)
    for c in prim.children:
        _anchor(c, base)


class _Composer:
    def __init__(self):
        self.layers: Dict[str, Stage] = {}
        self.stack: List[Tuple[str, str]] = []  # (layer file, prim path) being grafted: cycle detection

    def layer(self, path: str) -> Stage:
        """One layer file with its sublayers folded in (the layer stack), arcs anchored but unresolved. Cached."""
        path = os.path.abspath(path)
        if path in self.layers:
            return self.layers[path]
        stage = load_layer(path)
        self.layers[path] = stage
        base = os.path.dirname(path)
        for r in stage.roots:
            _anchor(r, base)
        for sub in _as_list(stage.metadata.get("subLayers")):
            asset = sub.asset if isinstance(sub, Ref) else sub
            weak = self.layer(os.path.join(base, asset))
            for k, v in weak.metadata.items():
                if k != "subLayers":
                    stage.metadata.setdefault(k, v)
            by_name = {r.name: r for r in stage.roots}
            for r in weak.roots:
                if r.name in by_name:
                    _merge(by_name[r.name], r, r.path, r.path, arcs_too=True)
                else:
                    stage.roots.append(_clone(r, None, r.name, r.path, r.path))
        return stage

    def compose_prim(self, prim: Prim, inherited: Optional[Dict[str, str]] = None) -> None:
        """Turns `prim` into its composed form, strongest first: its own opinions, the selected variants, the reference
        targets, the payload targets; then its children. `inherited`: variant selections of the prims that reference this
        one — a selection is taken from the strongest site that authors it, which is how a referencing layer picks a variant
        of a referenced asset."""
        sel = dict(prim.metadata.get("variants") or {}) if isinstance(prim.metadata.get("variants"), dict) else {}
        sel.update(inherited or {})
        # V: the selected variant of every set (a variant body may add nested sets, selections of them, and arcs)
        done = set()
        while True:
            pending = [s for s in prim.variant_sets if s not in done]
            if not pending:
                break
            for set_name in pending:
                done.add(set_name)
                choice = sel.get(set_name)
                body = prim.variant_sets[set_name].get(choice) if choice is not None else None
                if body is not None:
                    _merge(prim, body, prim.path, prim.path, arcs_too=True)
                    for sn, ch in (body.metadata.get("variants") or {}).items():
                        sel.setdefault(sn, ch)
        prim.variant_sets = {}
        # R then P: every target is copied into this prim's namespace, composed there under this prim's selections, and
        # folded in as the weaker side
        for arc in _ARC_KEYS:
            keep = []
            for tgt in _as_list(prim.metadata.pop(arc, None)):
                if not (isinstance(tgt, Ref) and tgt.asset):
                    keep.append(tgt)  # internal arc: the importer's business (native instancing, usd_import.rs:166-188)
                    continue
                tstage = self.layer(tgt.asset)
                tname = "/" + (tgt.path.strip("/") if tgt.path else str(tstage.metadata.get("defaultPrim", "")))
                target = tstage.prim_at(tname)
                if target is None:
                    raise ValueError(f"usd: {arc} target <{tname}> not found in {tgt.asset}")
                key = (tgt.asset, target.path)
                if key in self.stack:
                    raise ValueError(f"usd: composition cycle through <{target.path}> of {tgt.asset}")
                self.stack.append(key)
                try:
                    graft = _clone(target, prim.parent, prim.name, target.path, prim.path)
                    self.compose_prim(graft, sel)
                finally:
                    self.stack.pop()
                _merge(prim, graft, prim.path, prim.path)
            if keep:
                prim.metadata[arc] = keep if len(keep) > 1 else keep[0]
        for c in list(prim.children):
            self.compose_prim(c)


def _as_list(v):
    if v is None:
        return []
    return list(v) if isinstance(v, (list, tuple)) else [v]


def load_composed(path: str) -> Stage:
    """The stage rooted at `path`: its layer stack with every variant selection, external reference and payload flattened in."""
    comp = _Composer()
    stage = comp.layer(path)
    for r in list(stage.roots):
        comp.stack.append((os.path.abspath(path), r.path))
        try:
            comp.compose_prim(r)
        finally:
            comp.stack.pop()
    return stage
