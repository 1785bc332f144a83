"""Per-kernel device-time table of ONE warm adaptation step (torch.profiler / CUPTI; cheap alternative to an ncu
launch list while iterating).  Usage: python tools/step_profile.py [--eval]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def main():
    import vitta_b200
    from vitta_b200 import synth
    from vitta_b200.corpus.basics import OnlineAdapter
    from vitta_b200.models.tanet_models.tanet import TSN
    from vitta_b200.utils.opts import default_args
    dev = torch.device("cuda:0")
    vitta_b200.set_fp32_exact()
    K, T, RES, n = bench.K_CLASSES, bench.T, bench.RES, bench.N_PER_GPU
    model = TSN(K, T, 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256, tam=True,
                non_local=False, partial_bn=False)
    sd = synth.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd)
    model = model.to(dev)
    import numpy as np
    names = [k[:-len(".running_mean")] for k in sd if k.endswith("running_mean") and sd[k].dim() == 1 and ".tam." not in k]
    src_m = [np.zeros(sd[nm + ".weight"].shape[0], np.float32) for nm in names]
    src_v = [np.ones(sd[nm + ".weight"].shape[0], np.float32) for nm in names]
    args = default_args(arch='tanet', clip_length=T, batch_size=n, n_augmented_views=1, if_pred_consistency=False,
                        num_classes=K, input_size=RES)
    ad = OnlineAdapter(model, args, (src_m, src_v))
    x = synth.tanet_loader_tensor(synth.synth_video(n, 1, T, RES, seed=200, tag="tta")).to(dev)
    for _ in range(3):
        ad.adapt(x)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ad.adapt(x)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("total device time %.2f ms over %d kernels" % (tot / 1e3, sum(r[1] for r in rows)))
    for t, c, k in rows[:40]:
        print("%9.1f us %5.1f%% x%4d  %s" % (t, 100 * t / tot, c, k[:110]))


if __name__ == "__main__":
    main()
