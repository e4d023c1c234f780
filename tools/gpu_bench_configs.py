"""Training-step throughput of the secondary BASELINE.json configs on one B200 (CUDA events, 10 steps after 5 warm-up):
configs[1] SimpleViT CIFAR-100 shape, configs[0] README ViT (B=8), ViT-L/16 (B=128) and ViT-H/14 inference."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V
dev = torch.device("cuda:0")

def bench(name, model, B, img, classes, train=True, steps=10):
    model = model.to(dev)
    x = torch.randn(B, 3, img, img, device=dev).to(torch.bfloat16)
    y = torch.randint(0, classes, (B,), device=dev)
    if train:
        opt = V.FusedAdamW(model.parameters(), lr=2e-4, weight_decay=0.01)
        def step():
            opt.zero_grad()
            loss = V.softmax_cross_entropy(model(x), y, 0.1)
            loss.backward()
            opt.step()
            return loss
    else:
        model.eval()
        def step():
            with torch.no_grad():
                return model(x).float().sum()
    for _ in range(5):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%-34s B=%4d  %8.2f ms/step  %9.0f img/s  (%s, last value %.4f)" % (name, B, ms, B / ms * 1e3, "train" if train else "inference", float(out)), flush=True)

which = sys.argv[1:] or ["simple", "readme", "l16"]
torch.manual_seed(0)
if "simple" in which:
    bench("SimpleViT CIFAR 32/4 d512 L6 h8", V.SimpleViT(image_size=32, patch_size=4, num_classes=100, dim=512, depth=6, heads=8, mlp_dim=2048), 1024, 32, 100)
if "readme" in which:
    bench("README ViT 256/32 d1024 L6 h16", V.ViT(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048), 8, 256, 1000)
if "l16" in which:
    m = V.vit_l_16()
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
    bench("ViT-L/16 224", m, 128, 224, 1000)
if "robust" in which:   # robust=True: softmax + 3 Sinkhorn iterations in every attention layer (utils.py:1025-1037)
    bench("SimpleViT CIFAR robust=True", V.SimpleViT(image_size=32, patch_size=4, num_classes=100, dim=512, depth=6, heads=8, mlp_dim=2048, robust=True), 1024, 32, 100)
    for rb in (False, True):
        m = V.vit_b_16(robust=rb)
        with torch.no_grad():
            m.heads.head.weight.normal_(std=0.02)
        bench("ViT-B/16 224 robust=%s" % rb, m, 256, 224, 1000)
        del m
    bench("ViT-H/14 224 robust=True inference", V.vit_h_14(robust=True), 16, 224, 1000, train=False)
if "h14" in which:
    bench("ViT-H/14 224 inference", V.vit_h_14(), 64, 224, 1000, train=False)
