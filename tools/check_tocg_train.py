import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"oracle")); sys.path.insert(0, os.path.join(ROOT,"tests"))
torch.set_num_threads(16)
import hrv_loader; hrv_loader.load()
import hrviton_oracle as orc
from helpers import synth_state_dict, tocg_opt
from hrviton_b200 import synth, autograd_tocg as at, autograd_g as ag
import networks
n, h, w, seed = 2, 128, 96, 11
sd = synth_state_dict("tocg", seed)
i1, i2 = synth.tocg_inputs(n, h, w, seed)
m = networks.ConditionGenerator(tocg_opt(True), 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d); m.load_state_dict(sd); m = m.cuda().train()
def rel(a, b): return float((a.float().cpu() - b.float()).norm() / (b.float().norm() + 1e-12))
with torch.no_grad():
    a_ref, b_ref = i1, i2
    a = ag.FromNCHW.apply(i1.cuda(), None, None); b = ag.FromNCHW.apply(i2.cuda(), None, None)
    for k in range(5):
        a_ref = orc.resblock(sd, "ClothEncoder.%d" % k, a_ref, "down", True)
        b_ref = orc.resblock(sd, "PoseEncoder.%d" % k, b_ref, "down", True)
        a = at._resblock(m.ClothEncoder[k], a); b = at._resblock(m.PoseEncoder[k], b)
        print("enc level %d: cloth rel %.3e pose rel %.3e  (shape %s)" % (k, rel(a.permute(0,3,1,2), a_ref), rel(b.permute(0,3,1,2), b_ref), tuple(a_ref.shape)))
    x_ref = orc.resblock(sd, "conv", b_ref, "same", True); x = at._resblock(m.conv, b)
    print("conv(same) rel %.3e" % rel(x.permute(0,3,1,2), x_ref))
    x_ref = orc.resblock(sd, "SegDecoder.0", x_ref, "up", True); x = at._resblock(m.SegDecoder[0], x)
    print("SegDecoder.0(up) rel %.3e" % rel(x.permute(0,3,1,2), x_ref))
    # eval-mode comparison of the same blocks for reference (product inference path)
    m.eval()
    fl, seg, wc, wcm = m(i1.cuda(), i2.cuda())
    fr, sr, _, _ = orc.tocg_forward(sd, i1, i2)
    print("eval-mode seg rel %.3e flow4 rel %.3e" % (rel(seg, sr), rel(fl[-1], fr[-1])))
    fr, sr, _, _ = orc.tocg_forward(sd, i1, i2, bn_train=True)
    m.train()
    fl, seg, wc, wcm = at.tocg_forward_train(m, i1.cuda(), i2.cuda())
    print("train-mode seg rel %.3e flow4 rel %.3e  flows rel %s" % (rel(seg, sr), rel(fl[-1], fr[-1]), [round(rel(a_, b_), 4) for a_, b_ in zip(fl, fr)]))
