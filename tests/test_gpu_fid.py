"""sinFID (src/sinFID/): the block-0 feature networks on the GPU against their oracle restatements, weight loading, and
north_star's "sinFID within a stated delta": SVFID of GPU-generated samples vs SVFID of ORACLE-generated samples (same
weights, same z, same refinement noise, same real clip, same feature weights)."""
import numpy as np
import pytest
import torch

from oracle import hpvg_oracle as orc
from util import rel_l2

pytestmark = pytest.mark.gpu

# stated deltas on the mean per-sample Fréchet distance (relative to the oracle's value)
SVFID_DELTA = {"bf16": 2e-2, "tf32": 1e-3}      # measured: 1.7e-3 / 6.9e-3 (bf16, 5 / 10 scales), 6e-5 (tf32)


def _ncdhw(feat_cl_tensor, ops):
    return ops.unpack_cl(feat_cl_tensor).numpy()


def test_c3d_block0_matches_oracle(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg import fid, ops
    feats = fid.C3DBlock0(seed=3)
    w, b = feats.weights()
    assert w.shape == (64, 3, 3, 3, 3) and not feats.pretrained
    x = np.tanh(np.random.default_rng(0).standard_normal((2, 3, 5, 30, 41))).astype(np.float32)
    ref = orc.c3d_block0(x, w, b).numpy()
    got = _ncdhw(feats(hp.from_numpy(x)), ops)
    assert got.shape == ref.shape == (2, 64, 5, 30, 41)
    assert rel_l2(got, ref) < 1e-2
    # inputs in (0, 1) with the reference's normalisation (c3d.py:129-130)
    f01 = fid.C3DBlock0(seed=3, normalize_input=True)
    x01 = (x + 1) / 2
    assert rel_l2(_ncdhw(f01(hp.from_numpy(x01)), ops), orc.c3d_block0(x01, w, b, normalize_input=True).numpy()) < 1e-2


@pytest.mark.parametrize("hw", [(65, 81), (48, 64), (192, 257)])
def test_inception_block0_matches_oracle(hpvg_gpu, hw):
    hp = hpvg_gpu
    from hpvg import fid, ops
    feats = fid.InceptionBlock0(seed=5)
    x = np.tanh(np.random.default_rng(1).standard_normal((2, 3) + hw)).astype(np.float32)
    ref = orc.inception_block0(x, feats.params).numpy()
    got = _ncdhw(feats(hp.from_numpy(x)), ops)[:, :, 0]
    assert ref.shape[2:] == fid.InceptionBlock0.out_hw(*hw)
    assert got.shape == ref.shape
    e = rel_l2(got, ref)
    assert e < 1.5e-2, "inception block 0 (3 bf16 conv layers) rel-L2 %.3e" % e


def test_feature_weights_load_from_npz(hpvg_gpu, tmp_path):
    hp = hpvg_gpu
    from hpvg import fid, ops
    rng = np.random.default_rng(2)
    # C3D: MindSpore names
    w, b = (rng.standard_normal((64, 3, 3, 3, 3)) * 0.1).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    np.savez(tmp_path / "c3d.npz", **{"conv1.weight": w, "conv1.bias": b})
    f = fid.C3DBlock0(weights=str(tmp_path / "c3d.npz"))
    assert f.pretrained and np.array_equal(f.weights()[0], w) and np.array_equal(f.weights()[1], b)
    # Inception: torchvision names -> same features as the same numbers under MindSpore-hub names
    a = fid.InceptionBlock0(seed=9)
    tv = {}
    for name, _, _ in fid.InceptionBlock0.SPEC:
        tv[name + "_3x3.conv.weight"] = a.params[name + ".conv.weight"]
        tv[name + "_3x3.bn.weight"] = a.params[name + ".bn.gamma"]
        tv[name + "_3x3.bn.bias"] = a.params[name + ".bn.beta"]
        tv[name + "_3x3.bn.running_mean"] = a.params[name + ".bn.moving_mean"]
        tv[name + "_3x3.bn.running_var"] = a.params[name + ".bn.moving_variance"]
    np.savez(tmp_path / "inc.npz", **tv)
    bnet = fid.InceptionBlock0(weights=str(tmp_path / "inc.npz"))
    x = hp.from_numpy(np.tanh(rng.standard_normal((1, 3, 40, 52))).astype(np.float32))
    assert np.array_equal(ops.unpack_cl(a(x)).numpy(), ops.unpack_cl(bnet(x)).numpy())
    with pytest.raises(hp.HpvgError):
        fid.C3DBlock0(weights={"nothing": w})


def _generator(hp, n_body, seed):
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    params = orc.randomize_bn_stats(orc.init_generator_params(oopt, n_body, seed=seed), opt=oopt)
    net = n3.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        net.init_next_stage()
    net.load_parameters(params)
    return net, opt, oopt, params


@pytest.mark.parametrize("mode,n_body,n_samples", [("bf16", 4, 16), ("tf32", 4, 16), ("bf16", 9, 2)])
def test_svfid_of_gpu_samples_vs_oracle_samples(hpvg_gpu, mode, n_body, n_samples):
    """The number north_star asks for.  n_samples clips are generated twice from the SAME weights, z and refinement
    noise — by the CUDA path and by the fp32 oracle — and each set is scored against the same real clip with the same
    C3D block-0 weights: GPU clips -> GPU features -> device moments -> Fréchet; oracle clips -> torch-CPU features ->
    np.mean / np.cov -> Fréchet (fid_score.py:160-178, 219-242).  16 samples at a 5-scale pyramid in both precision
    modes, 2 samples at the full 10-scale pyramid (13 x 192 x 257).  The delta is asserted and printed."""
    hp = hpvg_gpu
    from hpvg import fid
    hp.set_precision(mode)
    try:
        net, opt, oopt, params = _generator(hp, n_body, seed=31)
        rng = np.random.default_rng(17)
        amps = [1.0, 0.0, 0.0] + [0.35] * (n_body - 2)
        top = orc.scale_shape(oopt, n_body)
        real = np.tanh(rng.standard_normal((1, 3) + top)).astype(np.float32)
        feats = fid.C3DBlock0(seed=77)
        w, b = feats.weights()
        pt = orc.to_torch(params)
        gpu_rows, ref_clips, clip_err = [], [], []
        bs = 4 if n_body <= 4 else 1
        for c0 in range(0, n_samples, bs):
            z = rng.standard_normal((bs, 128) + orc.scale_shape(oopt, 0)).astype(np.float32)
            nz = {s: rng.standard_normal((bs, 3) + orc.scale_shape(oopt, s)).astype(np.float32)
                  for s in range(opt.vae_levels, n_body + 1)}
            tz = hp.from_numpy(z)
            x, _ = net(tz, amps, noise_init=tz, isRandom=True, noises={k: hp.from_numpy(v) for k, v in nz.items()})
            gpu_rows.append(fid.sample_moments(feats(x)).numpy())
            with torch.no_grad():
                rx, _ = orc.generator_forward(None, amps, pt, oopt, noise_init=torch.from_numpy(z), is_random=True,
                                              noises={k: torch.from_numpy(v) for k, v in nz.items()})
            ref_clips.append(rx.numpy())
            clip_err.append(rel_l2(x.numpy(), rx.numpy()))
        gpu_rows, ref_clips = np.concatenate(gpu_rows), np.concatenate(ref_clips)
        count = int(np.prod(top))
        real_row = fid.sample_moments(feats(hp.from_numpy(real))).numpy()[0]
        got, per_gpu = fid.svfid_from_moments(real_row, gpu_rows, count)
        with torch.no_grad():
            want, per_ref = orc.svfid(real, ref_clips, lambda c: orc.c3d_block0(c, w, b).numpy())
        delta = abs(got - want) / abs(want)
        worst = max(abs(a - r) / abs(r) for a, r in zip(per_gpu, per_ref))
        print("SVFID %s, %d-scale pyramid %s, %d samples: GPU %.6f vs oracle %.6f  (relative delta %.3e, worst single "
              "sample %.3e; clips rel-L2 %.3e)" % (mode, n_body + 1, top, n_samples, got, want, delta, worst,
                                                   max(clip_err)))
        assert delta < SVFID_DELTA[mode], (got, want)
    finally:
        hp.set_precision("bf16")
