"""CPU: the Python drop-in surface has the reference's call signatures.

tests/golden/api.json is recorded from the UNMODIFIED reference by ``python -m oracle.make_api_golden``: for every hook /
loss / meter / model / driver entry point that ``vitta_b200`` mirrors under the same module path, the parameter names in
order, their kinds and their defaults.  A mirror may add keyword parameters AFTER the reference's (all with defaults), never
rename, reorder or re-default one (a required parameter may become optional) -- a caller written against the reference must keep working unchanged."""
import importlib
import inspect
import json
import os

import pytest

import cases

API = json.load(open(os.path.join(cases.GOLDEN_DIR, "api.json")))


def _default(v):
    if v is inspect.Parameter.empty:
        return "<required>"
    if isinstance(v, (int, float, str, bool, type(None))):
        return v
    if isinstance(v, (list, tuple)):
        return [_default(x) for x in v]
    return "<%s>" % getattr(v, "__name__", type(v).__name__)


def _describe(obj):
    fn = obj.__init__ if inspect.isclass(obj) else obj
    return [[n, p.kind.name, _default(p.default)] for n, p in inspect.signature(fn).parameters.items() if n != "self"]


@pytest.mark.parametrize("key", sorted(API))
def test_signature_matches_reference(key):
    modname, dotted = key.split(":")
    obj = importlib.import_module("vitta_b200." + modname)
    for part in dotted.split("."):
        assert hasattr(obj, part), "vitta_b200.%s has no %s" % (modname, dotted)
        obj = getattr(obj, part)
    want, got = API[key], _describe(obj)
    if any(k == "VAR_KEYWORD" for _, k, _ in want):          # f(*args, **kwargs) in the reference: anything goes
        return
    assert len(got) >= len(want), (key, got, want)
    for (wn, wk, wd), (gn, gk, gd) in zip(want, got):
        assert gn == wn, "%s: parameter %r where the reference has %r" % (key, gn, wn)
        assert gk == wk or gk == "POSITIONAL_OR_KEYWORD", (key, gn, gk, wk)
        # a parameter the reference requires may be optional here (callers written against the reference pass it anyway)
        assert gd == wd or wd == "<required>", "%s: default of %r is %r, reference %r" % (key, gn, gd, wd)
    for gn, gk, gd in got[len(want):]:                        # extensions must be optional
        assert gd != "<required>" or gk in ("VAR_POSITIONAL", "VAR_KEYWORD"), (key, gn)
