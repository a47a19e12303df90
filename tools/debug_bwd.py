"""GPU diagnostic: per-stage backward of the PatchGAN against autograd intermediates of the oracle."""
import os, sys, json
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200
from sscg_b200.arch import define_Dis
from oracle import ref_arch as RA

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / max(b.norm().item(), 1e-30))

def main(prec, cin, H):
    torch.manual_seed(0)
    net = define_Dis(cin, 64, "n_layers", norm="instance", gpu_ids=[0])
    net.precision = prec
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, cin, H, H) * 2 - 1
    # oracle with intermediates
    xr = x.clone().requires_grad_(True)
    raws, acts = [], []
    h = F.conv2d(xr, sd["dis_model.0.weight"], sd["dis_model.0.bias"], 2, 1); raws.append(h)
    h = F.leaky_relu(h, 0.2); acts.append(h)
    for idx, st in ((2, 2), (3, 2), (4, 1)):
        r = F.conv2d(h, sd["dis_model.%d.0.weight" % idx], sd["dis_model.%d.0.bias" % idx], st, 1); raws.append(r)
        h = F.leaky_relu(F.instance_norm(r, eps=1e-5), 0.2); acts.append(h)
    out = F.conv2d(h, sd["dis_model.5.weight"], sd["dis_model.5.bias"], 1, 1); raws.append(out)
    for t in raws + acts:
        t.retain_grad()
    probe = torch.randn(out.shape)
    (out * probe).sum().backward()
    # fused
    xg = x.cuda()
    runner = net._runner
    runner._setup(xg.device, prec)
    runner.ensure_weights()
    plan = runner.plan(2, H, H)
    c = plan.acquire_ctx()
    plan.forward(c, xg.contiguous())
    y = plan.output_nchw(c)
    print("fwd", rel(y, out))
    # forward intermediates
    for i in range(4):
        a = c.act[i + 1].as_nhwc_f32().permute(0, 3, 1, 2)
        print(" act", i + 1, rel(a[:, :acts[i].shape[1]], acts[i]))
    def hook(i, pl):
        torch.cuda.synchronize()
        _, _, ho, wo = pl.geom[i]
        cp = pl.weights[i].Co_pitch
        d = pl.draw[: 2 * ho * wo * cp].view(2, ho, wo, cp).float()
        if pl.draw_lo is not None:
            d = d + pl.draw_lo[: 2 * ho * wo * cp].view(2, ho, wo, cp).float()
        d = d.permute(0, 3, 1, 2)
        g = raws[i].grad
        print(" stage", i, "draw", rel(d[:, :g.shape[1]], g), "maxabs", float(g.abs().max()))
        if i > 0:
            ga = pl.gact[i].as_nhwc_f32().permute(0, 3, 1, 2)
            print("          gact", rel(ga[:, :acts[i - 1].shape[1]], acts[i - 1].grad))
    plan.backward(c, probe.cuda(), need_dx=True, need_dw=True, debug_hook=hook)

if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for prec in ("bf16x3", "bf16"):
        for cin, H in ((3, 64), (21, 128), (3, 128)):
            print("=====", prec, cin, H)
            main(prec, cin, H)
