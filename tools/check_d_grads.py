import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"oracle")); sys.path.insert(0, os.path.join(ROOT,"tests"))
import hrv_loader; hrv_loader.load()
import hrviton_oracle as orc
from helpers import gen_opt, synth_state_dict
from hrviton_b200 import synth, autograd_g
import network_generator
n, h, w, seed = 2, 128, 96, 31
sd = synth_state_dict("gend", seed)
x, seg = synth.gen_inputs(n, h, w, seed, input_nc=3)
inp = torch.cat([seg, x], 1)
for which in ["all", "d0_last", "d1_last", "d0_f0", "d0_f1"]:
    sdr = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("weight_u", "weight_v"))) for k, v in sd.items()}
    inp_ref = inp.clone().requires_grad_(True)
    res_ref = orc.gen_d_forward(sdr, inp_ref)
    def L(res):
        if which == "all": return sum((f * (1 + 0.1 * j)).mean() for fs in res for j, f in enumerate(fs))
        if which == "d0_last": return res[0][-1].mean()
        if which == "d1_last": return res[1][-1].mean()
        if which == "d0_f0": return res[0][0].mean()
        if which == "d0_f1": return res[0][1].mean()
    L(res_ref).backward()
    m = network_generator.MultiscaleDiscriminator(gen_opt(h, w, True)); m.load_state_dict(sd); m = m.cuda().eval()
    inp_d = inp.cuda().requires_grad_(True)
    res = autograd_g.discriminator_forward_train(m, inp_d, need_wgrad=True)
    L(res).backward()
    g, gr = inp_d.grad.cpu(), inp_ref.grad
    print(which, "input grad rel %.3e  (|ref| %.3e |got| %.3e)" % (float((g-gr).norm()/gr.norm()), float(gr.norm()), float(g.norm())))
    for name, p in m.named_parameters():
        gr = sdr[name].grad
        if gr is None or p.grad is None or float(gr.norm()) < 1e-9: continue
        print("   %-40s rel %.3e" % (name, float((p.grad.float().cpu()-gr).norm()/gr.norm())))
