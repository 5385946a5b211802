import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtgs_b200.ssim import ssim
dev = torch.device("cuda:0")
H, W = 1080, 1920
gt = torch.rand(1, 3, H, W, device=dev)
pred = (gt * 0.8 + 0.1 * torch.rand(1, 3, H, W, device=dev)).requires_grad_(True)
m = torch.rand(H, W, 1, device=dev) < 0.9
for _ in range(3):
    pred.grad = None
    (1 - ssim(gt, pred, data_range=1.0, mask=m)).backward()
torch.cuda.synchronize()
