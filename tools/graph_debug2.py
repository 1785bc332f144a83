import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import cases
import vitta_b200
from vitta_b200 import synth
from vitta_b200.corpus.basics import OnlineAdapter
from vitta_b200.models.tanet_models.tanet import TSN
from vitta_b200.utils.opts import default_args
dev = torch.device("cuda:0")
vitta_b200.set_fp32_exact()
cfg = cases.TANET_CASES["tanet_t8_r64_stats_mse"]
g = cases.load_golden("tanet_t8_r64_stats_mse")
src_m, src_v = cases.src_stats_from_golden(g)
model = TSN(cfg["K"], cfg["T"], 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256, tam=True,
            non_local=False, partial_bn=False)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=1))
model.base_model.fc.p = 0.0
model = model.to(dev)
args = default_args(arch='tanet', clip_length=cfg["T"], batch_size=cfg["N"], n_augmented_views=1, if_pred_consistency=False,
                    reg_type="l1_loss", lr=1e-3, num_classes=cfg["K"], input_size=cfg["res"], moving_avg=True, cuda_graph=True)
ad = OnlineAdapter(model, args, (src_m, src_v))
x = synth.tanet_loader_tensor(synth.synth_video(cfg["N"], 1, cfg["T"], cfg["res"], seed=500, tag="tta")).to(dev)
torch.autograd.set_detect_anomaly(True, check_nan=False)
for s in range(6):
    r = ad.adapt(x)
    print("step", s, float(r["loss_reg"]), "graph" if ad._graph is not None else "eager")
