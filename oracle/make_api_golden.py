"""Record the Python call signatures of the reference's drop-in surface from the UNMODIFIED
reference as a fixture: parameter names in order and the defaults of every hook / loss / model / driver entry point that
``vitta_b200`` mirrors under the same module path.  Run in the build container only (/root/reference is not on the GPU box):

    python -m oracle.make_api_golden      ->  tests/golden/api.json
"""
import importlib
import inspect
import json
import os
import sys

from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SURFACE = {
    "utils.norm_stats_utils": ["CombineNormStatsRegHook_onereg", "ComputeNormStatsHook", "compute_regularization",
                               "compute_kld", "CombineNormStatsRegHook_onereg.hook_fn",
                               "CombineNormStatsRegHook_onereg.add_hook_back", "CombineNormStatsRegHook_onereg.close"],
    "utils.BNS_utils": ["BNFeatureHook", "choose_layers", "freeze_except_bn", "collect_bn_params"],
    "utils.pred_consistency_utils": ["compute_pred_consis"],
    "utils.utils_": ["AverageMeter", "AverageMeter.update", "AverageMeterTensor", "AverageMeterTensor.update",
                     "MovingAverageTensor", "MovingAverageTensor.update", "accuracy"],
    "models.tanet_models.tanet": ["TSN", "TSN.forward"],
    "models.tanet_models.temporal_module": ["TAM", "TemporalBottleneck"],
    "models.tanet_models.basic_ops": ["ConsensusModule"],
    "models.videoswintransformer_models.recognizer3d": ["Recognizer3D", "Recognizer3D.forward"],
    "models.videoswintransformer_models.swin_transformer": ["SwinTransformer3D", "WindowAttention3D",
                                                            "SwinTransformerBlock3D", "PatchMerging", "PatchEmbed3D",
                                                            "BasicLayer"],
    "models.videoswintransformer_models.i3d_head": ["I3DHead"],
    "corpus.basics": ["tta_standard", "compute_statistics", "validate", "get_model", "get_dataset_tanet",
                      "get_dataset_videoswin"],
    "corpus.main_eval": ["eval"],
}


def _default(v):
    if v is inspect.Parameter.empty:
        return "<required>"
    if isinstance(v, (int, float, str, bool, type(None))):
        return v
    if isinstance(v, (list, tuple)):
        return [_default(x) for x in v]
    return "<%s>" % getattr(v, "__name__", type(v).__name__)      # classes / functions (e.g. nn.LayerNorm): by name


def describe(obj):
    fn = obj.__init__ if inspect.isclass(obj) else obj
    rows = []
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self":
            continue
        rows.append([name, str(p.kind.name), _default(p.default)])
    return rows


def resolve(mod, dotted):
    obj = mod
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def main():
    ref_harness.load_reference()
    out = {}
    for modname, names in SURFACE.items():
        mod = importlib.import_module(modname)
        for n in names:
            out["%s:%s" % (modname, n)] = describe(resolve(mod, n))
    path = os.path.join(ROOT, "tests", "golden", "api.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out), "signatures")


if __name__ == "__main__":
    sys.exit(main())
