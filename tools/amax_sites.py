"""Which tensors of one warm TANet adaptation step still get a standalone operand-range pass (vitta_amax_f32), i.e. have no
producer kernel that emits max|x| on the way: shape and the Python call site.  Usage: python tools/amax_sites.py"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def main():
    import vitta_b200
    from vitta_b200 import ops
    dev = torch.device("cuda:0")
    from vitta_b200 import synth
    vitta_b200.set_fp32_exact()
    ad = bench.build_tanet(dev, None, bench.N_PER_GPU, graph=False)[0]
    x = synth.tanet_loader_tensor(synth.synth_video(bench.N_PER_GPU, 1, bench.T, bench.RES, seed=200, tag="tta")).to(dev)
    for _ in range(2):
        ad.adapt(x)
    torch.cuda.synchronize()
    real = ops.amax_f32

    def spy(t, out=None):
        fr = [f for f in traceback.extract_stack()[:-1] if "vitta_b200" in f.filename][-3:]
        print("amax_f32", tuple(t.shape), " <- ", " / ".join("%s:%d %s" % (os.path.basename(f.filename), f.lineno, f.name) for f in fr))
        return real(t, out)
    ops.amax_f32 = spy
    print("== TANet step")
    ad.adapt(x)
    torch.cuda.synchronize()
    if "--swin" in sys.argv:
        ops.amax_f32 = real
        del ad, x
        torch.cuda.empty_cache()
        ad = bench.build_swin(dev, None, "tiny", 8)
        x = synth.swin_loader_tensor(synth.synth_video(8, 2, 32, 224, seed=200, tag="tta")).to(dev)
        for _ in range(2):
            ad.adapt(x)
        torch.cuda.synchronize()
        from vitta_b200 import ops_swin
        ops.amax_f32 = spy
        if getattr(ops_swin, "amax_f32", None) is real:
            ops_swin.amax_f32 = spy
        print("== Video-Swin-T step")
        ad.adapt(x)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
