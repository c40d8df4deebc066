"""SURVEY.md §8e/§8f rows on the GPU: device-side sinFID moments, sharded sample generation (shard invariance),
checkpoint round trip with the reference's parameter names, the progressive training driver, and the NCCL
communicator (world size 1 in-process; N > 1 runs under bench.py / tools/fid_multi_gpu.py on a multi-GPU box)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc
from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu


def _small_generator(hp, n_body=3, img_size=64, seed=11):
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(img_size=img_size), orc.default_opt(img_size=img_size)
    params = orc.randomize_bn_stats(orc.init_generator_params(oopt, n_body, seed=seed), opt=oopt)
    net = n3.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        net.init_next_stage()
    net.load_parameters(params)
    return net, opt, oopt, params


def test_sample_moments_match_numpy_mean_and_cov(hpvg_gpu):
    """fid_score.py:160-178: mu = mean over positions, sigma = np.cov(rowvar=False) of the 64-channel feature map."""
    hp = hpvg_gpu
    from hpvg import fid, ops
    rng = np.random.default_rng(0)
    f = bf16_round(rng.standard_normal((2, 64, 3, 20, 70)) * 0.7 + 0.1)
    rows = fid.sample_moments(ops.pack_cl(hp.from_numpy(f))).numpy()
    count = 3 * 20 * 70
    for n in range(2):
        act = f[n].reshape(64, -1).T.astype(np.float64)
        mu, sigma = fid.moments_to_stats(rows[n], count)
        assert np.allclose(mu, act.mean(0), atol=1e-4)
        assert rel_l2(sigma, np.cov(act, rowvar=False)) < 2e-3


def test_svfid_gpu_matches_cpu_restatement_within_delta(hpvg_gpu):
    """sinFID 'within a stated delta' (north_star): the SAME feature weights on both sides, GPU clips + device moments
    vs oracle clips + numpy statistics.  Delta: 5 % relative (bf16 pyramid) on the mean Fréchet distance."""
    hp = hpvg_gpu
    from hpvg import fid, sampling
    net, opt, oopt, params = _small_generator(hp)
    amps = [1.0, 0.5, 0.4, 0.3]
    feats = fid.RandomFeatures3D(3, seed=5)
    w, b = feats.weights()
    rng = np.random.default_rng(9)
    real = np.tanh(rng.standard_normal((1, 3) + orc.scale_shape(oopt, 3))).astype(np.float32)
    n_samples = 4
    # GPU: sharded path (single process) with host noise injected for the refinement levels through a fixed device seed
    rows, count, clips = sampling.generate_moments(net, amps, n_samples, feats, batch=2, seed=3, keep=True)
    real_rows = fid.sample_moments(feats(hp.from_numpy(real))).numpy()
    got, per = fid.svfid_from_moments(real_rows[0], rows, count)

    def cpu_stats(clip):
        act = F.leaky_relu(F.conv3d(torch.from_numpy(bf16_round(clip)), torch.from_numpy(bf16_round(w)),
                                    torch.from_numpy(b), padding=1), 0.2)[0].reshape(64, -1).T.numpy().astype(np.float64)
        return act.mean(0), np.cov(act, rowvar=False)
    m1, s1 = cpu_stats(real)
    # the clips themselves are covered by the generator parity tests; here the statistics pipeline is under test, so the
    # CPU side consumes the SAME clips
    want = float(np.mean([fid.calculate_frechet_distance(m1, s1, *cpu_stats(clips[i:i + 1])) for i in range(n_samples)]))
    assert abs(got - want) <= 0.05 * abs(want) + 1e-3, (got, want)


def test_generation_is_shard_invariant(hpvg_gpu):
    """Sample i is bit-identical whatever the (rank, world, batch) partition: every noise source is keyed by the
    global sample index (host z via default_rng([seed, i]); device Philox via (seed, sample index, element))."""
    hp = hpvg_gpu
    from hpvg import sampling
    net, opt, oopt, params = _small_generator(hp)
    amps = [1.0, 0.5, 0.4, 0.3]
    idx_all, ref = sampling.generate(net, amps, 6, rank=0, world=1, batch=6, seed=7)
    got = {}
    for rank in range(2):
        idx, clips = sampling.generate(net, amps, 6, rank=rank, world=2, batch=2, seed=7)
        for i, c in zip(idx, clips):
            got[i] = c
    assert sorted(got) == list(range(6))
    for i in range(6):
        assert np.array_equal(got[i], ref[idx_all.index(i)]), "sample %d depends on the partition" % i


def test_checkpoint_roundtrip_with_reference_names(hpvg_gpu, tmp_path):
    hp = hpvg_gpu
    from hpvg import checkpoint as ck, networks_3d as n3
    net, opt, oopt, params = _small_generator(hp, n_body=2)
    f = ck.save_checkpoint(net, str(tmp_path / "netG_2.ckpt"))
    assert f.endswith(".npz")
    state = ck.load_checkpoint(f)
    assert set(state) == set(params)
    for k in ("decoder.0.1.bn2d.moving_mean", "body.1.6.weight", "encode._features.0.0.weight_u"):
        assert np.array_equal(state[k], params[k])
    net2 = n3.GeneratorHPVAEGAN(opt, seed=99)
    for _ in range(2):
        net2.init_next_stage()
    assert ck.load_param_into_net(net2, state) == []
    z = hp.from_numpy(np.random.default_rng(1).standard_normal((1, 128) + orc.scale_shape(oopt, 0)).astype(np.float32))
    a = net(z, [1, 0, 0], noise_init=z)[0].numpy()
    b = net2(z, [1, 0, 0], noise_init=z)[0].numpy()
    assert np.array_equal(a, b)
    # a MindSpore-format .ckpt with the reference's parameter names (body.0.0.N. for the later stages) loads as it is
    ms_path = ck.save_mindspore_ckpt(ck.to_reference_names(state), str(tmp_path / "netG_ref.ckpt"))
    ref_named = ck.load_checkpoint(ms_path)
    assert any(k.startswith("body.0.0.1.") for k in ref_named) and not any(k.startswith("body.1.") for k in ref_named)
    net3 = n3.GeneratorHPVAEGAN(opt, seed=77)
    for _ in range(len(net.body)):
        net3.init_next_stage()
    assert ck.load_param_into_net(net3, ref_named) == []
    assert np.array_equal(a, net3(z, [1, 0, 0], noise_init=z)[0].numpy())
    ck.save_json({"noise_amps": [1, 0.25], "scale_idx": 1}, str(tmp_path / "intermediate.json"))
    assert ck.load_json(str(tmp_path / "intermediate.json"))["scale_idx"] == 1
    with pytest.raises(hp.HpvgError):
        ck.load_param_into_net(net2, {"decoder.0.0.weight": state["decoder.0.0.weight"]})


@pytest.mark.parametrize("graph,niter", [(False, 2), (True, 5)])
def test_progressive_training_driver_vae_then_gan(hpvg_gpu, tmp_path, graph, niter):
    """train_video.py:413-419 over scales 0..3 of a small pyramid (vae_levels=3 -> scales 0-2 VAE phase, scale 3 GAN
    phase with a fresh D): parameter groups / lr schedule, noise-amp calibration, per-scale state files — launched
    kernel by kernel, and with every scale's iteration captured as a CUDA graph after two eager iterations."""
    hp = hpvg_gpu
    from hpvg import driver, networks_3d as n3, checkpoint as ck
    from hpvg.utils import images as uimg
    opt = uimg.default_opt(img_size=64)
    np.random.seed(0)
    rng = np.random.default_rng(0)
    clips = {s: np.tanh(rng.standard_normal((1, 3) + uimg.scale_shape(opt, s))).astype(np.float32) for s in range(4)}
    G = n3.GeneratorHPVAEGAN(opt, seed=2)
    # parameter groups (train_video.py:76-105)
    groups, body_idx, codec = driver.generator_param_groups(opt, G, 0)
    assert codec and body_idx == () and abs(groups[0]["lr"] - opt.lr_g) < 1e-12
    seen = []
    before = G.parameters_dict()["decoder.6.weight"].numpy().copy()
    amps, hist = driver.train_pyramid(opt, G, n3.WDiscriminator3D, lambda s: clips[s], niter=niter, stop_scale=3,
                                      save_dir=str(tmp_path), on_iter=lambda s, it, l: seen.append((s, it)), graph=graph)
    assert seen == [(s, it) for s in range(4) for it in range(niter)]
    assert all(np.isfinite(float(l[1])) for h in hist for l in h)
    assert not np.array_equal(G.parameters_dict()["decoder.6.weight"].numpy(), before)          # the VAE phase trained it
    if graph:       # replayed iterations keep training: the loss of the VAE phase at scale 0 goes down over 5 iterations
        assert float(hist[0][-1][1]) < float(hist[0][0][1])
    assert len(G.body) == 3 and len(amps) == 4 and amps[0] == 1
    assert all(a > 0 for a in amps[1:])                       # noise_amp_init * RMSE(real, reconstruction)
    groups, body_idx, codec = driver.generator_param_groups(opt, G, 3)
    assert not codec and body_idx == (2,) and abs(groups[0]["lr"] - opt.lr_g) < 1e-12
    G2 = n3.GeneratorHPVAEGAN(opt, seed=2)
    G2.init_next_stage()
    G2.init_next_stage()
    groups, body_idx, codec = driver.generator_param_groups(opt, G2, 2)      # VAE phase, scale 2: codec at lr*0.2^2
    assert codec and body_idx == (1,) and abs(groups[0]["lr"] - opt.lr_g * opt.lr_scale ** 2) < 1e-15
    assert abs(groups[2]["lr"] - opt.lr_g) < 1e-15 and len(groups) == 3
    for s in range(3):
        assert hist[s][0][0] is None and np.isfinite(float(hist[s][1][1]))       # VAE phase: G loss only
    assert np.isfinite(float(hist[3][1][0])) and np.isfinite(float(hist[3][1][1]))
    inter = json.load(open(tmp_path / "intermediate.json"))
    assert inter["scale_idx"] == 3 and len(inter["noise_amps"]) == 4
    assert os.path.exists(tmp_path / "netG_3.npz") and os.path.exists(tmp_path / "netD_3.npz")
    assert not os.path.exists(tmp_path / "netD_2.npz")
    state = ck.load_checkpoint(str(tmp_path / "netG_3"))
    assert "body.2.6.weight" in state and np.isfinite(state["body.2.0.0.weight"]).all()


def test_nccl_communicator_world1(hpvg_gpu):
    """ncclCommInitRank / ncclAllGather through ctypes on our stream (world 1: gather == identity)."""
    hp = hpvg_gpu
    from hpvg import dist
    comm = dist.NcclCommunicator(0, 1, key="pytest%d" % os.getpid())
    st = hp.Stream()
    rows = hp.from_numpy(np.arange(12, dtype=np.float32).reshape(3, 4))
    out = comm.all_gather_rows(rows, stream=st)
    assert np.array_equal(out, np.arange(12, dtype=np.float32).reshape(3, 4))
    comm.close()


def test_host_uniform_noise_becomes_standard_normal_on_the_device(hpvg_gpu):
    """SamplePipeline's z: uniforms drawn on the host per sample, Box-Muller on the device (ops.box_muller_): N(0,1)
    statistics, no infinities at u = 0, the documented pairing, and reproducible per (seed, index)."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    from hpvg import sampling
    u = sampling.host_uniform_for_sample(11, 5, (128, 4, 24, 33))
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert np.array_equal(u, sampling.host_uniform_for_sample(11, 5, (128, 4, 24, 33)))
    z = ops.box_muller_(hp.from_numpy(u)).numpy()
    assert np.isfinite(z).all() and abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 3)) < 2e-2 and abs(np.mean(z ** 4) - 3.0) < 5e-2
    f = u.ravel().astype(np.float64)
    r = np.sqrt(-2.0 * np.log(1.0 - f[0::2]))
    ref = np.stack([r * np.cos(2 * np.pi * f[1::2]), r * np.sin(2 * np.pi * f[1::2])], axis=1).ravel()
    assert np.max(np.abs(z.ravel() - ref)) < 1e-4
    edge = ops.box_muller_(hp.from_numpy(np.array([0.0, 0.0, 0.99999994, 0.25, 0.5], np.float32))).numpy()
    assert np.isfinite(edge).all() and edge[0] == 0.0 and abs(edge[4]) < 1e-6 + 2.0   # odd tail element pairs with u2 = 0.5
