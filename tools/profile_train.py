"""torch.profiler view of one stage-2 train step (which kernels — ours vs library/torch glue — take the time)."""
import os, sys, types, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRV_VGG_RANDOM_INIT", "1")
import bench
import hrv_loader; hrv_loader.load()
import network_generator, networks
from hrviton_b200 import train_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(batch, 1024, 768)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
