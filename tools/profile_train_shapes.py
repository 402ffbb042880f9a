"""Which torch (non-hrv) ops dominate one stage-2 step, grouped by input shapes."""
import os, sys, types, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("HRV_VGG_RANDOM_INIT", "1")
import bench
import hrv_loader; hrv_loader.load()
import network_generator, networks
from hrviton_b200 import train_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
topt = types.SimpleNamespace(warp_feature="T1", out_layer="relu", cuda=True)
tocg = networks.ConditionGenerator(topt, 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d).to(dev).eval()
g = bench.build_generator(dev).train()
dopt = bench.gen_opt(); dopt.ndf, dopt.norm_D, dopt.n_layers_D, dopt.num_D, dopt.no_ganFeat_loss = 64, "spectralinstance", 3, 2, False
D = network_generator.MultiscaleDiscriminator(dopt); D.init_weights("xavier", 0.02); D = D.to(dev).train()
vgg = networks.Vgg19().to(dev).eval()
tr = train_step.Stage2Trainer(tocg, g, D, vgg)
batch = train_step.synthetic_batch(B, 1024, 768, dev, seed=1)
for _ in range(2): tr.step(batch, 1024, 768)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    tr.step(batch, 1024, 768)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if not e.key.startswith("hrv") and e.self_device_time_total > 150]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print("%-38s %8.3f ms  x%-4d %s" % (e.key[:38], e.self_device_time_total / 1e3, e.count, str(e.input_shapes)[:110]))
