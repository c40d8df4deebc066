import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpvg
from hpvg import networks_3d as n3, ops, train as T
from hpvg.utils import images as uimg
from oracle import hpvg_oracle as orc
from util import rel_l2, bf16_round
hpvg.init(0)
opt, oopt = uimg.default_opt(), orc.default_opt()
pg = orc.init_generator_params(oopt, 1, seed=3)
rng = np.random.default_rng(3)
for k in pg:
    if k.endswith("bias") or k.endswith("beta"):
        pg[k] = (rng.standard_normal(pg[k].shape) * 0.05).astype(np.float32)
G = n3.GeneratorHPVAEGAN(opt); G.init_next_stage(); G.load_parameters(pg); G.set_train(True)
shape = (1, 4, 30, 41)
x3 = bf16_round(rng.standard_normal((1, 3) + shape[1:]) * 0.5)
up = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32) * 0.3
gout = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32)
# oracle
tg = orc.to_torch(pg, requires_grad=("body.",))
taps = {}
xt = torch.from_numpy(x3).requires_grad_(True)
with orc.bf16_emulation():
    pre = orc.block_forward(xt, tg, "body.0.", oopt, True, taps=taps)
    out = torch.tanh(pre + torch.from_numpy(up))
for k, v in taps.items():
    if v.requires_grad: v.retain_grad()
out.backward(torch.from_numpy(gout))
# gpu
ws = n3.Workspace()
xin = ops.pack_cl(hpvg.from_numpy(x3), c_pitch=8)
xw = ops.pack_cl(hpvg.from_numpy(x3), c_pitch=64, zero_to=64)
o, ctxs = T.block_forward_train(G.body[0], xin, hpvg.from_numpy(up), ws, "s0", x_wide=xw)
print("fwd out", rel_l2(o.numpy(), out.detach().numpy()))
for j in range(6):
    print(" fwd layer", j, "a", rel_l2(ops.unpack_cl(ctxs[j]["a"]).numpy(), taps["body.0.%d.out" % j].detach().numpy()),
          "y", rel_l2(ops.unpack_cl(ctxs[j]["y"]).numpy(), bf16_round(taps["body.0.%d.conv" % j].detach().numpy())))
book = T.GradBook()
g_pre, dx = T.block_backward(G.body[0], ctxs, hpvg.from_numpy(gout), book, ws, "s0", need_dx=True)
print("g_pre", rel_l2(g_pre.numpy(), taps["body.0.6.conv"].grad.numpy()))
for j in range(5, -1, -1):
    gy = ops.unpack_cl(ws.get("s0.%d.gy" % j, ctxs[j]["a"].shape, hpvg.BF16)).numpy()
    print(" bwd layer", j, "gy(conv out grad)", rel_l2(gy, taps["body.0.%d.conv" % j].grad.numpy()),
          "ga(act grad)" , "-" if j == 5 else rel_l2(ops.unpack_cl(ws.get("s0.%d.dx" % (j + 1), ctxs[j]["a"].shape, hpvg.BF16)).numpy(), taps["body.0.%d.out" % j].grad.numpy()))
ga5 = ops.unpack_cl(ws.get("s0.t.dx", ctxs[5]["a"].shape, hpvg.BF16)).numpy()
print(" ga5 (tail dgrad)", rel_l2(ga5, taps["body.0.5.out"].grad.numpy()))
print("dx", rel_l2(dx.numpy(), xt.grad.numpy()))
pd = G.parameters_dict()
for k, t in tg.items():
    if t.grad is not None and np.linalg.norm(t.grad.numpy()) > 1e-5:
        print(k, rel_l2(book.of(pd[k]).numpy(), t.grad.numpy()))
