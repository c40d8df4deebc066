import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpvg as hp
from hpvg import train as T
from oracle import hpvg_oracle as orc
from test_gpu_train import _setup
hp.init(0)
def run(mode, iters):
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 3, seed=5)
    G.noise_seed = 0x1234567
    s0, s3 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, 3)
    st = hp.Stream()
    real = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + s3)).astype(np.float32))
    real_zero = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32))
    noise = hp.from_numpy(rng.standard_normal((1, 128) + s0).astype(np.float32))
    amps = [1.0, 0.0, 0.0, 0.3]
    block = G.body[-1]
    optG = T.ClippedAdam(opt, [{"params": T.trainable_params(block), "lr": opt.lr_g}], opt.lr_g, beta1=0.5, beta2=0.999, device_step=True)
    optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=0.5, beta2=0.999, device_step=True)
    g_step = T.TrainOneStepCell(T.GWithLoss(opt, D, G, device_rng=True), optG, cells_to_invalidate=[block])
    d_step = T.TrainOneStepCell(T.DWithLoss(opt, D, G, alpha=0.37, device_rng=True), optD, cells_to_invalidate=[D])
    g_step.set_train(); d_step.set_train()
    it = T.GraphedIteration(st, g_step, d_step, real, real_zero, noise, amps, dict(isVAE=False, trainable_body=(2,)))
    it.warmup(1)
    losses = []
    if mode == "graph":
        it.capture()
        for _ in range(iters): losses.append(it())
    else:
        for _ in range(iters): losses.append(it._body(True))
    st.sync()
    state = {"D." + k: t.numpy() for k, t in D.parameters_dict().items()}
    state.update({"G." + k: t.numpy() for k, t in G.parameters_dict().items()})
    state["G.draws"] = g_step.network.trainer.draws.numpy(); state["D.draws"] = d_step.network.trainer.draws.numpy()
    state["G.step"] = optG.d_step.numpy(); state["D.step"] = optD.d_step.numpy()
    return state, [(float(a), float(b)) for a, b in losses]
for iters in (1, 2):
    a, la = run("eager", iters); b, lb = run("eager", iters); c, lc = run("graph", iters)
    print("iters", iters, "eager", la, "eager2", lb, "graph", lc)
    for k in a:
        e1 = np.abs(a[k].astype(np.float64) - b[k]).max(); e2 = np.abs(a[k].astype(np.float64) - c[k]).max()
        if e2 > 1e-6 or e1 > 1e-6: print("  %-40s eager-vs-eager %.3e  eager-vs-graph %.3e  |x| %.3e" % (k, e1, e2, np.abs(a[k]).max()))
