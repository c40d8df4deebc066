"""Development diagnostic: per-layer teacher-forced backward of one refinement block against the oracle at a given
size (default: the finest scale), with the intermediate quantities that located the BatchNorm-backward rounding bias
(DESIGN.md §5).  python tools/debug_tf_full.py [N,T,H,W]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import hpvg as hp
from oracle import hpvg_oracle as orc
from util import rel_l2, bf16_round
import test_gpu_train as tt
from hpvg import networks_3d as n3, ops, train as T
hp.init(0)
shape = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (1, 13, 192, 257)
G, D, opt, oopt, pg, pd, rng = tt._setup(hp, 1)
G.set_train(True)
x3 = bf16_round(rng.standard_normal((1, 3) + shape[1:]) * 0.5)
up = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32) * 0.3
gout = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32)
tg = orc.to_torch(pg, requires_grad=("body.",))
taps = {}
xt = torch.from_numpy(x3).requires_grad_(True)
with orc.bf16_emulation():
    pre = orc.block_forward(xt, tg, "body.0.", oopt, True, taps=taps)
    out = torch.tanh(pre + torch.from_numpy(up))
for v in taps.values():
    if v.requires_grad:
        v.retain_grad()
out.backward(torch.from_numpy(gout))
print(sorted(taps.keys())[:12])
block = G.body[0]
pdict = G.parameters_dict()
ws = n3.Workspace()
for j in range(opt.num_layer + 1):
    layer = block.layers[j]
    xin = x3 if j == 0 else taps["body.0.%d.out" % (j - 1)].detach().numpy()
    x_cl = ops.pack_cl(hp.from_numpy(xin), c_pitch=8 if j == 0 else 64)
    a, ctx = T.layer_forward_train(layer, x_cl, ws, "tf%d" % j)
    if j == 0:
        ctx["x_wide"] = x_cl
    a_ref = taps["body.0.%d.out" % j].detach().numpy()
    e_a = rel_l2(ops.unpack_cl(a).numpy(), a_ref)
    ga = taps["body.0.%d.out" % j].grad.numpy()
    book = T.GradBook()
    dx = T.layer_backward(layer, ctx, ops.pack_cl(hp.from_numpy(ga)), book, ws, "tf%d" % j, need_dx=True)
    pre_n = "body.0.%d." % j
    errs = {nm: rel_l2(book.of(pdict[pre_n + nm]).numpy(), tg[pre_n + nm].grad.numpy()) for nm in ("0.weight", "1.bn2d.gamma", "1.bn2d.beta")}
    ref_dx = xt.grad.numpy() if j == 0 else taps["body.0.%d.out" % (j - 1)].grad.numpy()
    got_dx = dx.numpy() if j == 0 else ops.unpack_cl(dx).numpy()
    # conv-output gradient if the oracle tapped it
    extra = ""
    kc = "body.0.%d.conv" % j
    if kc in taps and taps[kc].grad is not None:
        gy_ref = taps[kc].grad.numpy()
        gy = ops.unpack_cl(ws.get("tf%d.gy" % j, a.shape, hp.BF16)).numpy()
        extra = " gy %.3e |gy_ref| %.3e |dw_ref| %.3e" % (rel_l2(gy, gy_ref), np.linalg.norm(gy_ref), np.linalg.norm(tg[pre_n + "0.weight"].grad.numpy()))
    print("layer %d: act %.3e  dW %.3e dgamma %.3e dbeta %.3e dx %.3e%s" % (j, e_a, errs["0.weight"], errs["1.bn2d.gamma"], errs["1.bn2d.beta"], rel_l2(got_dx, ref_dx), extra))
    if j == 1 and extra:
        d = (gy.astype(np.float64) - gy_ref.astype(np.float64))
        N = gy.shape[2] * gy.shape[3] * gy.shape[4]
        print("   sum over voxels per channel: ours max|.| %.3e  ref max|.| %.3e ; diff sum max %.3e ; rms(gy_ref) %.3e ; N %d" % (
            np.abs(gy.sum((0, 2, 3, 4), dtype=np.float64)).max(), np.abs(gy_ref.sum((0, 2, 3, 4), dtype=np.float64)).max(),
            np.abs(d.sum((0, 2, 3, 4))).max(), float(np.sqrt((gy_ref.astype(np.float64) ** 2).mean())), N))
        xin64 = xin.astype(np.float64)
        print("   mean(x) per channel (max) %.3e rms(x) %.3e" % (np.abs(xin64.mean((0, 2, 3, 4))).max(), np.sqrt((xin64**2).mean())))
        # wgrad of OUR kernel on the ORACLE's fp32->bf16 gy vs torch
        dw_k = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
        ops.conv_wgrad_cl(x_cl, ops.pack_cl(hp.from_numpy(gy_ref)), dw_k)
        print("   wgrad kernel on bf16(gy_ref): rel-L2 vs oracle dW %.3e" % rel_l2(dw_k.numpy(), tg[pre_n + "0.weight"].grad.numpy()))
        import torch.nn.functional as F
        w0 = torch.zeros(64, 64, 3, 3, 3, requires_grad=True)
        F.conv3d(torch.from_numpy(xin), w0, None, padding=1).backward(torch.from_numpy(bf16_round(gy_ref)))
        print("   torch wgrad on bf16(gy_ref): rel-L2 vs oracle dW %.3e" % rel_l2(w0.grad.numpy(), tg[pre_n + "0.weight"].grad.numpy()))
        w1 = torch.zeros(64, 64, 3, 3, 3, requires_grad=True)
        F.conv3d(torch.from_numpy(xin), w1, None, padding=1).backward(torch.from_numpy(gy))
        print("   torch wgrad on OUR gy: rel-L2 vs oracle dW %.3e ; vs our dW %.3e" % (rel_l2(w1.grad.numpy(), tg[pre_n + "0.weight"].grad.numpy()), rel_l2(w1.grad.numpy(), book.of(pdict[pre_n + "0.weight"]).numpy())))
    if j == 1:
        yv = ops.unpack_cl(ctx["y"]).numpy().astype(np.float64)      # stored bf16 conv output
        sv = ctx["saved"].numpy().astype(np.float64)                  # scale, shift, mean, invstd
        mu_true = yv.mean((0, 2, 3, 4)); var_true = yv.var((0, 2, 3, 4))
        print("   saved mean vs mean(stored y): max abs diff %.3e (|mean| max %.3e) ; invstd rel diff max %.3e" % (
            np.abs(sv[2] - mu_true).max(), np.abs(mu_true).max(), np.abs(sv[3] * np.sqrt(var_true + 1e-5) - 1).max()))
        gz = ga.astype(np.float64) * np.where(ops.unpack_cl(a).numpy() > 0, 1.0, 0.2)
        xh = (yv - sv[2][None, :, None, None, None]) * sv[3][None, :, None, None, None]
        print("   sum xh per channel max %.3e ; m0 %.3e m1 %.3e (max abs)" % (np.abs(xh.sum((0, 2, 3, 4))).max(),
              np.abs(gz.mean((0, 2, 3, 4))).max(), np.abs((gz * xh).mean((0, 2, 3, 4))).max()))
        sc32 = sv[0].astype(np.float32)[None, :, None, None, None]
        ga_b = bf16_round(ga).astype(np.float32)
        mask = np.where(ops.unpack_cl(a).numpy() > 0, np.float32(1.0), np.float32(0.2))
        gz32 = ga_b * mask
        m0 = gz32.astype(np.float64).mean((0, 2, 3, 4)).astype(np.float32)[None, :, None, None, None]
        m1 = (gz32.astype(np.float64) * xh).mean((0, 2, 3, 4)).astype(np.float32)[None, :, None, None, None]
        r32 = sc32 * (gz32 - m0 - xh.astype(np.float32) * m1)
        print("   simulated fp32 gy: sum max %.3e ; after bf16 RN: sum max %.3e ; GPU gy sum max %.3e ; sim-vs-GPU rel %.3e" % (
            np.abs(r32.sum((0, 2, 3, 4), dtype=np.float64)).max(), np.abs(bf16_round(r32).sum((0, 2, 3, 4), dtype=np.float64)).max(),
            np.abs(gy.sum((0, 2, 3, 4), dtype=np.float64)).max(), rel_l2(gy, bf16_round(r32))))
        print("   scale (gamma*invstd) range %.4f .. %.4f ; mean|ga| %.3e" % (sv[0].min(), sv[0].max(), np.abs(ga).mean()))
        gsum = gy.sum((0, 2, 3, 4), dtype=np.float64); c = int(np.abs(gsum).argmax())
        print("   worst channel %d: GPU sum %.3e, sc*N*m0 %.3e, sc %.4f" % (c, gsum[c], sv[0][c] * N * float(m0.ravel()[c]), sv[0][c]))
