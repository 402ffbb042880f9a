"""One stage-2 training step (frozen tocg -> warp -> SPADE G -> D; hinge + feature-matching + VGG; Adam), restating the
body of the reference loop train_generator.py:201-360 as a callable so bench.py / tests can drive it on synthetic batches.

Hot path = the three networks (kernels of this repo).  The per-step tensor glue between them (nearest/bilinear resizes,
15x15 Gaussian blur, argmax -> one-hot -> 7-class regroup, hi-res grid_sample of the cloth) is SURVEY.md §8(f) row N2
("next"): it runs as torch ops here."""
import torch
import torch.nn.functional as F

LABELS7 = [[0], [2, 4, 7, 8, 9, 10, 11], [3], [1], [5], [6], [12]]  # train_generator.py:261-269


def gaussian_blur_15_3(x):
    """tgm.image.GaussianBlur((15,15),(3,3)) restated: depth-wise separable 15-tap Gaussian, sigma 3, zero padding 7
    (torchgeometry is absent; parity of this glue op is unpinned)."""
    k = torch.arange(15, dtype=torch.float32, device=x.device) - 7
    g = torch.exp(-(k * k) / (2 * 3.0 * 3.0))
    g = g / g.sum()
    c = x.shape[1]
    x = F.conv2d(x, g.view(1, 1, 1, 15).expand(c, 1, 1, 15), padding=(0, 7), groups=c)
    return F.conv2d(x, g.view(1, 1, 15, 1).expand(c, 1, 15, 1), padding=(7, 0), groups=c)


def make_generator_inputs(tocg, batch, fine_h, fine_w, occlusion=False):
    """train_generator.py:201-275 (opt.GT False, clothmask_composition 'warp_grad')."""
    from .tocg import make_grid
    cm, c_paired = batch["cloth_mask"], batch["cloth"]
    with torch.no_grad():
        pre_cm_down = F.interpolate(cm, size=(256, 192), mode="nearest")
        parse_agn_down = F.interpolate(batch["parse_agnostic"], size=(256, 192), mode="nearest")
        clothes_down = F.interpolate(c_paired, size=(256, 192), mode="bilinear")
        dense_down = F.interpolate(batch["densepose"], size=(256, 192), mode="bilinear")
        input1 = torch.cat([clothes_down, pre_cm_down], 1)
        input2 = torch.cat([parse_agn_down, dense_down], 1)
        flow_list, fake_segmap, _, warped_cm = tocg(input1, input2)
        mask = torch.ones_like(fake_segmap)
        mask[:, 3:4] = warped_cm
        fake_segmap = fake_segmap * mask
        n, _, ih, iw = c_paired.shape
        grid = make_grid(n, ih, iw).to(c_paired.device)
        flow = F.interpolate(flow_list[-1].permute(0, 3, 1, 2), size=(ih, iw), mode="bilinear").permute(0, 2, 3, 1)
        flow_norm = torch.cat([flow[..., 0:1] / ((96 - 1.0) / 2.0), flow[..., 1:2] / ((128 - 1.0) / 2.0)], 3)
        warped_grid = grid + flow_norm
        warped_cloth = F.grid_sample(c_paired, warped_grid, padding_mode="border", align_corners=False)
        warped_clothmask = F.grid_sample(cm, warped_grid, padding_mode="border", align_corners=False)
        fake_parse_gauss = gaussian_blur_15_3(F.interpolate(fake_segmap, size=(ih, iw), mode="bilinear"))
        fake_parse = fake_parse_gauss.argmax(dim=1)[:, None]
        if occlusion:
            so = F.softmax(fake_parse_gauss, dim=1)
            warped_clothmask = warped_clothmask - torch.cat([so[:, 1:3], so[:, 5:]], 1).sum(1, keepdim=True) * warped_clothmask
            warped_cloth = warped_cloth * warped_clothmask + (1 - warped_clothmask)
        old_parse = torch.zeros(n, 13, fine_h, fine_w, device=cm.device).scatter_(1, fake_parse, 1.0)
        parse = torch.stack([old_parse[:, idx].sum(1) for idx in LABELS7], 1)
        g_in = torch.cat((batch["agnostic"], batch["densepose"], warped_cloth), 1)
    return g_in.detach(), parse.detach()


class Stage2Trainer:
    """Holds the optimisers / criteria of train_generator.py:145-159 and runs one step."""

    def __init__(self, tocg, generator, discriminator, vgg, lambda_feat=10.0, lambda_vgg=10.0, g_lr=1e-4, d_lr=4e-4,
                 reducers=None):
        from .spade import GANLoss
        self.tocg, self.G, self.D, self.vgg = tocg, generator, discriminator, vgg
        self.crit_gan = GANLoss("hinge")
        self.vgg_weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
        self.lambda_feat, self.lambda_vgg = lambda_feat, lambda_vgg
        self.opt_g = torch.optim.Adam(generator.parameters(), lr=g_lr, betas=(0.0, 0.9), fused=True)
        self.opt_d = torch.optim.Adam(discriminator.parameters(), lr=d_lr, betas=(0.0, 0.9), fused=True)
        self.reducers = reducers or {}  # {"G": GradBucketReducer, "D": ...} for data-parallel runs

    @staticmethod
    def _split(pred):
        fake = [[t[:t.size(0) // 2] for t in p] for p in pred]
        real = [[t[t.size(0) // 2:] for t in p] for p in pred]
        return fake, real

    def step(self, batch, fine_h, fine_w):
        from . import autograd_g
        g_in, parse = make_generator_inputs(self.tocg, batch, fine_h, fine_w)
        im = batch["image"]
        # ---------------- generator update (train_generator.py:279-322)
        out = self.G(g_in, parse)
        d_in = torch.cat((torch.cat((parse, out), 1), torch.cat((parse, im), 1)), 0)
        pred = autograd_g.discriminator_forward_train(self.D, d_in, need_wgrad=False, as_float=False)  # D grads are zeroed before use (:354)
        pred_fake, pred_real = self._split(pred)
        loss_gan = self.crit_gan(pred_fake, True, for_discriminator=False)
        loss_feat = 0
        num_d = len(pred_fake)
        for i in range(num_d):
            for j in range(len(pred_fake[i]) - 1):
                loss_feat = loss_feat + (pred_fake[i][j] - pred_real[i][j].detach()).abs().mean(dtype=torch.float32) * self.lambda_feat / num_d
        loss_vgg = autograd_g.vgg_loss(self.vgg, self.vgg_weights, out, im) * self.lambda_vgg
        loss_gen = (loss_gan + loss_feat + loss_vgg).mean()
        self.opt_g.zero_grad(set_to_none=True)
        loss_gen.backward()
        if "G" in self.reducers:
            self.reducers["G"].reduce()
        self.opt_g.step()
        # ---------------- discriminator update (train_generator.py:327-360)
        with torch.no_grad():
            out2 = self.G(g_in, parse)
        d_in = torch.cat((torch.cat((parse, out2), 1), torch.cat((parse, im), 1)), 0)
        pred = autograd_g.discriminator_forward_train(self.D, d_in, need_wgrad=True, as_float=False)
        pred_fake, pred_real = self._split(pred)
        loss_dis = (self.crit_gan(pred_fake, False, for_discriminator=True) + self.crit_gan(pred_real, True, for_discriminator=True)).mean()
        self.opt_d.zero_grad(set_to_none=True)
        loss_dis.backward()
        if "D" in self.reducers:
            self.reducers["D"].reduce()
        self.opt_d.step()
        return {"loss_gen": loss_gen.detach(), "loss_dis": loss_dis.detach(), "gan": loss_gan.detach(),
                "feat": loss_feat.detach() if torch.is_tensor(loss_feat) else loss_feat, "vgg": loss_vgg.detach()}


def synthetic_batch(n, h, w, device, seed=0):
    """Synthetic VITON-HD-shaped batch (shapes/ranges of cp_dataset.py outputs)."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def smooth(c):
        t = torch.rand((n, c, h // 16, w // 16), generator=g) * 2 - 1
        return F.interpolate(t, size=(h, w), mode="bilinear", align_corners=False)

    def onehot(c):
        lab = torch.randint(0, c, (n, h // 32, w // 32), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2)
        return torch.zeros(n, c, h, w).scatter_(1, lab[:, None], 1.0)

    b = {"cloth": smooth(3), "cloth_mask": (smooth(1) > 0).float(), "parse_agnostic": onehot(13), "densepose": smooth(3),
         "agnostic": smooth(3), "image": smooth(3)}
    return {k: v.to(device) for k, v in b.items()}
