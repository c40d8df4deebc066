import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpvg
from hpvg import networks_3d as n3, ops
from hpvg.utils import images as uimg
from oracle import hpvg_oracle as orc
from util import rel_l2
hpvg.init(0)
g = np.load(os.path.join(ROOT, "tests/golden/sample_small.npz"))
kw = {"img_size": int(g["img_size"])}
opt, oopt = uimg.default_opt(**kw), orc.default_opt(**kw)
nb = int(g["n_body"])
params = orc.randomize_bn_stats(orc.init_generator_params(oopt, nb, seed=int(g["seed"])), opt=oopt)
net = n3.GeneratorHPVAEGAN(opt)
for _ in range(nb): net.init_next_stage()
net.load_parameters(params)
pt = orc.to_torch(params)
noises = {int(k[6:]): g[k] for k in g.files if k.startswith("noise_")}
taps = {}
with torch.no_grad():
    rx, rv = orc.generator_forward(None, list(g["amps"]), pt, oopt, noise_init=torch.from_numpy(g["z"]), is_random=True,
                                   noises={k: torch.from_numpy(v) for k, v in noises.items()}, taps=taps)
z = hpvg.from_numpy(g["z"])
hn = {k: hpvg.from_numpy(v) for k, v in noises.items()}
x, vae = net(z, list(g["amps"]), noise_init=z, isRandom=True, noises=hn)
print("vae", rel_l2(vae.numpy(), rv.numpy()))
for idx in range(nb):
    o = net.ws.get("out%d" % idx, (2, 3) + uimg.scale_shape(opt, idx + 1), hpvg.F32).numpy()
    print("end-to-end stage", idx, rel_l2(o, taps["body.%d.out" % idx].numpy()))
# teacher forcing: feed the oracle's previous output
prev = rv.numpy()
for idx in range(nb):
    xin = hpvg.from_numpy(prev)
    # run only stage idx
    size = uimg.scale_shape(opt, idx + 1)
    add = opt.vae_levels <= idx + 1
    up, xi = ops.upsample_noise_pack(xin, size, noise=hn.get(idx + 1) if add else None, amp=float(g["amps"][idx + 1]) if add else 0.0)
    out = net._run_block(net.body[idx], xi, up, "tf%d" % idx, None)
    ref = taps["body.%d.out" % idx].numpy()
    print("teacher-forced stage", idx, rel_l2(out.numpy(), ref), " in:", rel_l2(ops.unpack_cl(xi, C=3).numpy(), taps["body.%d.in" % idx].numpy()))
    # layer by layer inside the stage with oracle inputs
    h_ref = taps["body.%d.in" % idx]
    for j in range(6):
        lay = net.body[idx].layers[j]
        inp = ops.pack_cl(hpvg.from_numpy(h_ref.numpy()), c_pitch=8 if j == 0 else 64)
        y = ops.unpack_cl(lay.forward_cl(inp)).numpy()
        r = taps["body.%d.%d.out" % (idx, j)].numpy()
        print("   layer", j, "rel", rel_l2(y, r), "ref rms", float(np.sqrt((r**2).mean())))
        h_ref = taps["body.%d.%d.out" % (idx, j)]
    prev = ref
