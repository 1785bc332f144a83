"""CPU: the option surface is the reference's (utils/opts.py), pinned by tests/golden/opts.json which
oracle/make_opts_golden.py records from the unmodified reference parser: same destinations, option strings, types,
defaults, choices and flag kinds -- so scripts written against the reference's opts drive this package unchanged.
Only the author's site-specific path defaults are not mirrored (checked for presence and type only)."""
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "opts.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def _ours():
    from vitta_b200.utils import opts
    return {a.dest: a for a in opts.parser._actions if a.dest != "help"}, opts


def test_every_reference_option_exists_with_same_flags_type_and_default(gold):
    ours, _ = _ours()
    assert sorted(ours) == sorted(o["dest"] for o in gold["options"])
    for o in gold["options"]:
        a = ours[o["dest"]]
        assert sorted(a.option_strings) == o["flags"], o["dest"]
        assert type(a).__name__ == o["action"], o["dest"]
        assert (None if a.type is None else a.type.__name__) == o["type"], o["dest"]
        assert (None if a.choices is None else list(a.choices)) == o["choices"], o["dest"]
        assert a.nargs == o["nargs"], o["dest"]
        if not o["site_path"]:
            d = list(a.default) if isinstance(a.default, tuple) else a.default
            assert d == o["default"], o["dest"]


def test_normalisation_constants(gold):
    _, opts = _ours()
    assert opts.input_mean == gold["input_mean"] and opts.input_std == gold["input_std"]
    assert opts.img_norm_cfg == gold["img_norm_cfg"]


def test_get_opts_derived_fields_and_bool_pitfall():
    _, opts = _ours()
    a = opts.get_opts([])
    assert a.evaluate_baselines is False and a.baseline == "source"      # reference get_opts(): not args.tta, 'source'
    # type=bool parses any non-empty string as True (kept: scripts set these from Python, as the reference's do)
    assert opts.get_opts(["--tta", "False"]).tta is True
    assert opts.get_opts(["--wd", "0.001"]).weight_decay == 0.001
    assert opts.get_opts(["-j", "3", "-p", "7"]).workers == 3
