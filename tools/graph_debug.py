"""Debug aid: capture pieces of the TANet adaptation step into CUDA graphs and report which piece breaks capture."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import cases
import vitta_b200
from vitta_b200 import synth, ops
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
                    reg_type="l1_loss", lr=1e-3, num_classes=cfg["K"], input_size=cfg["res"], moving_avg=True)
ad = OnlineAdapter(model, args, (src_m, src_v))
x = synth.tanet_loader_tensor(synth.synth_video(cfg["N"], 1, cfg["T"], cfg["res"], seed=500, tag="tta")).to(dev)
for _ in range(3):
    ad.adapt(x)
torch.cuda.synchronize()


def attempt(name, fn):
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(gr):
            fn()
        gr.replay()
        torch.cuda.synchronize()
        print("CAPTURE OK  ", name)
    except Exception as e:
        print("CAPTURE FAIL", name, "::", str(e).splitlines()[0])
        traceback.print_exc(limit=6)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass


m = ad.model
m.train()
for mod in m.modules():
    if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
        mod.eval()
xin = x.view(-1, 3, x.size(2), x.size(3)).view(cfg["N"], cfg["T"], 3, x.size(2), x.size(3))
with torch.no_grad():
    attempt("forward (no grad)", lambda: m(xin))
attempt("forward (grad, hooks)", lambda: m(xin))


def fwd_loss():
    out = m(xin)
    loss = sum(h.r_feature for h in ad.stat_reg_hooks)
    return loss


attempt("forward + loss", fwd_loss)


def fwd_bwd():
    loss = fwd_loss()
    ad.optimizer.zero_grad()
    loss.backward()


attempt("forward + backward", fwd_bwd)


def full():
    fwd_bwd()
    ad.optimizer.step()


attempt("forward + backward + sgd", full)
