import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "noise-robust-vit_b200"))
import vit_pytorch_robust as V
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = V.vit_h_14()
with torch.no_grad():
    m.heads.head.weight.normal_(std=0.02); m.class_token.normal_(std=0.02)
m = m.to(dev)
opt = V.FusedAdamW(m.parameters(), lr=2e-4, weight_decay=0.01)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16); lab = torch.randint(0, 1000, (B,), device=dev)
def step():
    opt.zero_grad(); loss = V.softmax_cross_entropy(m(V.add_gaussian_noise(img, 0.1)), lab, 0.1); loss.backward(); opt.step(); return loss
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("ViT-H/14 224 training B=%d: %.1f ms/step  %.0f img/s  (general tcgen05 attention forward / backward; loss %.4f)" % (B, ms, B / ms * 1e3, loss.item()))
