"""CPU: the source-statistics file format.  The reference writes two ``.npy`` files with
``np.save(list_of_vectors, allow_pickle=True)`` (corpus/basics.py:306-307) and reads them back with
``list(np.load(f, allow_pickle=True))`` (:482-483): a pickled 1-D object array with one (C,) float32 vector per norm
layer in ``named_modules()`` order.  Files written here must load through the reference's expression, files in the
reference's format must load here, and the ragged channel counts (64 ... 2048) must survive."""
import numpy as np

from vitta_b200.corpus import basics
from vitta_b200.utils.opts import default_args


def _vectors(seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(c).astype(np.float32) for c in (64, 64, 256, 512, 2048, 96)]


def test_saved_lists_load_through_the_reference_expression(tmp_path):
    means, variances = _vectors(0), [np.abs(v) for v in _vectors(1)]
    fm, fv = tmp_path / "list_spatiotemp_mean_x.npy", tmp_path / "list_spatiotemp_var_x.npy"
    basics.save_stat_list(str(fm), means)
    basics.save_stat_list(str(fv), variances)
    got_m = list(np.load(str(fm), allow_pickle=True))          # the reference's read, verbatim semantics
    got_v = list(np.load(str(fv), allow_pickle=True))
    assert len(got_m) == len(means)
    for a, b in zip(got_m + got_v, means + variances):
        assert a.dtype == np.float32 and a.shape == b.shape
        np.testing.assert_array_equal(a, b)


def test_reference_format_files_load_here(tmp_path):
    means, variances = _vectors(2), [np.abs(v) for v in _vectors(3)]
    # what np.save(list, allow_pickle=True) produced under the numpy the reference pins (ragged list -> object array)
    for name, vecs in (("m.npy", means), ("v.npy", variances)):
        arr = np.empty(len(vecs), dtype=object)
        arr[:] = vecs
        np.save(str(tmp_path / name), arr, allow_pickle=True)
    args = default_args(spatiotemp_mean_clean_file=str(tmp_path / "m.npy"),
                        spatiotemp_var_clean_file=str(tmp_path / "v.npy"))
    got_m, got_v = basics.load_source_statistics(args)
    assert isinstance(got_m, list) and len(got_m) == len(means)
    for a, b in zip(got_m + got_v, means + variances):
        np.testing.assert_array_equal(a, b)


def test_in_memory_statistics_bypass_the_files():
    means, variances = _vectors(4), _vectors(5)
    args = default_args(source_stats=(means, variances))
    got_m, got_v = basics.load_source_statistics(args)
    assert got_m[3] is means[3] and got_v[0] is variances[0]


def test_all_result_file_has_the_reference_layout(tmp_path):
    """utils/utils_.py:252-267 of the reference: path rule, 'name value' header of the public args, two separator lines,
    two blank lines; accuracies are appended after that by the entry scripts."""
    from vitta_b200.utils.opts import default_args
    from vitta_b200.utils.utils_ import get_writer_to_all_result
    args = default_args(arch="tanet", result_dir=str(tmp_path / "res" / "tta_gauss"))
    f = get_writer_to_all_result(args)
    f.write("12.5 13.0\n")
    f.close()
    import os
    files = os.listdir(args.result_dir)
    assert len(files) == 1 and files[0].endswith("_all_result") and len(files[0]) == len("20260101_000000_all_result")
    lines = open(os.path.join(args.result_dir, files[0])).read().split("\n")
    names = [a for a in dir(args) if a[0] != "_"]
    assert [ln.split(" ")[0] for ln in lines[:len(names)]] == names
    assert lines[names.index("arch")] == "arch tanet"
    assert lines[len(names):len(names) + 5] == ["#" * 29, "#" * 29, "", "", "12.5 13.0"]
    f = get_writer_to_all_result(args, custom_path=str(tmp_path / "custom"))
    f.close()
    (name,) = os.listdir(tmp_path / "custom")
    assert name.startswith(str(args.baseline) + "_") and name.endswith("_all_result")
