"""CPU: control flow of the two entry scripts (same names and flow as the reference's tta_tanet_ucf101.py /
tta_swin_ucf101.py) with ``eval`` stubbed out -- per-corruption loop, result directories, the all-result file in the
reference's layout, and that under a torchrun launch only rank 0 writes it."""
import os
import runpy
import sys

import pytest

import cases  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,arch", [("tta_tanet_ucf101.py", "tanet"), ("tta_swin_ucf101.py", "videoswintransformer")])
def test_entry_script_flow(script, arch, tmp_path, monkeypatch):
    import vitta_b200.corpus.main_eval as me
    seen = []

    def fake_eval(args=None, model=None):
        os.makedirs(args.result_dir, exist_ok=True)
        seen.append((args.arch, args.corruptions, args.val_vid_list, args.result_dir))
        return [12.3456 + len(seen)], None
    monkeypatch.setattr(me, "eval", fake_eval)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("VITTA_N_CORRUPTIONS", "3")
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.setattr(sys, "argv", [script])
    runpy.run_path(os.path.join(ROOT, script), run_name="__main__")
    assert [s[0] for s in seen] == [arch] * 3
    assert [s[1] for s in seen] == ["gauss_shuffled", "pepper_shuffled", "salt_shuffled"]
    assert all(c in v and c in d for _, c, v, d in seen)
    first_dir = seen[0][3]
    (name,) = [f for f in os.listdir(first_dir) if f.endswith("_all_result")]
    lines = open(os.path.join(first_dir, name)).read().split("\n")
    sep = lines.index("#" * 29)
    assert lines[sep + 1] == "#" * 29 and lines[sep + 2:sep + 4] == ["", ""]
    assert lines[sep + 4:sep + 7] == ["13.346", "14.346", "15.346"]          # one line of accuracies per corruption
    assert any(ln.startswith("arch " + arch) for ln in lines[:sep])
    # torchrun launch: ranks other than 0 adapt (eval is called) but do not write the results file
    before = sorted(os.listdir(first_dir))
    monkeypatch.setenv("RANK", "1")
    runpy.run_path(os.path.join(ROOT, script), run_name="__main__")
    assert len(seen) == 6 and sorted(os.listdir(first_dir)) == before
