"""Record the option surface of the UNMODIFIED reference (utils/opts.py) as a fixture: destination name, option strings,
type name, default, choices, nargs and action kind of every argument.  Run in the build container only
(/root/reference is not available on the GPU box):

    python oracle/make_opts_golden.py      ->  tests/golden/opts.json

Site-specific path defaults (the author's home directory) are recorded too but marked, so the parity test can skip their
values while still checking that the option exists with the same type."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VITTA_REFERENCE", "/root/reference")
SITE_PATHS = {"video_data_dir", "spatiotemp_mean_clean_file", "spatiotemp_var_clean_file", "val_vid_list", "result_dir",
              "model_path"}


def main():
    spec = importlib.util.spec_from_file_location("ref_opts", os.path.join(REF, "utils", "opts.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = []
    for act in mod.parser._actions:
        if act.dest == "help":
            continue
        kind = type(act).__name__
        default = act.default
        if isinstance(default, tuple):
            default = list(default)
        rows.append({"dest": act.dest, "flags": sorted(act.option_strings), "action": kind,
                     "type": None if act.type is None else act.type.__name__, "default": default,
                     "choices": None if act.choices is None else list(act.choices), "nargs": act.nargs,
                     "site_path": act.dest in SITE_PATHS})
    out = {"options": rows, "input_mean": mod.input_mean, "input_std": mod.input_std, "img_norm_cfg": mod.img_norm_cfg}
    path = os.path.join(ROOT, "tests", "golden", "opts.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(rows), "options")


if __name__ == "__main__":
    sys.exit(main())
