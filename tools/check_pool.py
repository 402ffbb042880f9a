import torch, torch.nn.functional as F
torch.manual_seed(0)
cur = torch.randn(2, 16, 12, 16, device="cuda").bfloat16().requires_grad_(True)
R = torch.randn(2, 10, 8, 6, device="cuda")
v = cur[..., :10].permute(0, 3, 1, 2)
y = F.avg_pool2d(v.float(), 3, stride=2, padding=1, count_include_pad=False)
(y * R).sum().backward(); g1 = cur.grad.clone(); cur.grad = None
v = cur[..., :10].permute(0, 3, 1, 2).float().contiguous()
y2 = F.avg_pool2d(v, 3, stride=2, padding=1, count_include_pad=False)
(y2 * R).sum().backward(); g2 = cur.grad.clone()
print("fwd diff", float((y - y2).abs().max()), "bwd diff", float((g1.float() - g2.float()).abs().max()), float(g2.float().abs().max()))
