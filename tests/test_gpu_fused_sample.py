"""hpvg_generator_sample (include/hpvg.h): the generator's random-mode forward as ONE C call must give the same bits as
the per-layer path `GeneratorHPVAEGAN.construct(noise_init=z, isRandom=True)` (reference networks_3d.py:406-451 as
driven by eval_video.py:53-82), for any batching of the samples, and stay within the sampling tolerance of the oracle."""
import numpy as np
import pytest

from oracle import hpvg_oracle as orc

pytestmark = pytest.mark.gpu


def _build(hp, opt_kw, n_body, seed):
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt = uimg.default_opt(**opt_kw)
    oopt = orc.default_opt(**opt_kw)
    params = orc.randomize_bn_stats(orc.init_generator_params(oopt, n_body, seed=seed), opt=oopt)
    net = n3.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        net.init_next_stage()
    net.load_parameters(params)
    net.set_train(False)
    return net, opt


@pytest.mark.parametrize("n_body", [0, 5])
def test_fused_entry_is_bit_identical_to_the_layer_path(hpvg_gpu, n_body):
    hp = hpvg_gpu
    from hpvg import sampling
    net, opt = _build(hp, {}, n_body, seed=4)
    amps = [1.0] + [0.0, 0.0] + [0.3, 0.25, 0.2, 0.15, 0.1, 0.1, 0.1][:max(n_body - 2, 0)]
    amps = (amps + [0.1] * 12)[:n_body + 1]
    rng = np.random.default_rng(7)
    zs = sampling.z_init_size(opt, 3)
    z = hp.from_numpy(rng.standard_normal(zs).astype(np.float32))
    st = hp.Stream()
    net(z, amps, noise_init=z, isRandom=True, stream=st)      # first call: packs the filter banks, folds BatchNorm
    net.sample_counter = 11
    l0 = hp.lib.hpvg_launch_count()
    x_ref, vae_ref = net(z, amps, noise_init=z, isRandom=True, stream=st)
    st.sync()
    n_layer_path = hp.lib.hpvg_launch_count() - l0
    x_ref, vae_ref = x_ref.numpy().copy(), vae_ref.numpy().copy()

    fused = sampling.FusedSampler(net, amps, batch=3, stream=st)
    vae = hp.Tensor(vae_ref.shape, hp.F32)
    l0 = hp.lib.hpvg_launch_count()
    x = fused(z, sample_base=11, vae_out=vae, stream=st)
    st.sync()
    n_fused = hp.lib.hpvg_launch_count() - l0
    assert x.numpy().shape == x_ref.shape
    assert np.array_equal(x.numpy(), x_ref), "clips differ between the fused entry and the layer path"
    assert np.array_equal(vae.numpy(), vae_ref)
    assert n_fused == n_layer_path, (n_fused, n_layer_path)      # the same launches, from one host call
    assert net.sample_counter == 14

    if n_body:
        # a sample's clip does not depend on how the samples are batched: sample 12 alone == row 1 of the batch
        z1 = hp.from_numpy(np.ascontiguousarray(z.numpy()[1:2]))
        x1 = fused(z1, sample_base=12, stream=st)
        st.sync()
        assert np.array_equal(x1.numpy()[0], x_ref[1])


def test_pipeline_with_the_fused_entry_gives_the_same_clips(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg import sampling
    net, opt = _build(hp, {}, 4, seed=5)
    amps = [1.0, 0.0, 0.0, 0.3, 0.2]
    got = {}
    for fused in (False, True):
        pipe = sampling.SamplePipeline(net, amps, batch=2, seed=3, fused=fused)
        clips = {}
        pipe.run(sampling.local_chunks(5, 2, 0, 1), sink=lambda chunk, a: clips.update({i: a[j].copy() for j, i in enumerate(chunk)}))
        pipe.close()
        got[fused] = clips
    assert sorted(got[True]) == [0, 1, 2, 3, 4]
    for i in range(5):
        assert np.array_equal(got[True][i], got[False][i]), "sample %d" % i
    idx, clips = sampling.generate(net, amps, 5, batch=2, seed=3, fused=True)      # the public loop, fused
    assert idx == [0, 1, 2, 3, 4]
    for i in range(5):
        assert np.array_equal(clips[i], got[False][i])


def test_fused_entry_rejects_bad_arguments(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg import sampling
    net, opt = _build(hp, {}, 2, seed=6)
    fused = sampling.FusedSampler(net, [1.0, 0.0, 0.0], batch=2)
    z = hp.Tensor(sampling.z_init_size(opt, 3), hp.F32).zero_()
    with pytest.raises(ValueError):
        fused(z)
    net.set_train(True)
    with pytest.raises(hp.HpvgError):
        sampling.FusedSampler(net, [1.0, 0.0, 0.0], batch=2)
    net.set_train(False)


def test_block_fwd_eval_is_bit_identical_to_the_layer_path(hpvg_gpu):
    """hpvg_block_fwd_eval: one refinement block (head 3->64, 3 x 64->64, tail 64->3 + residual, tanh) as one C call."""
    import ctypes
    hp = hpvg_gpu
    from hpvg import ops, sampling
    net, opt = _build(hp, {}, 4, seed=8)
    fused = sampling.FusedSampler(net, [1.0, 0.0, 0.0, 0.3, 0.2], batch=2)
    rng = np.random.default_rng(9)
    size = net.stage_shape(4)
    prev = hp.from_numpy(np.tanh(rng.standard_normal((2, 3) + tuple(net.stage_shape(3)))).astype(np.float32))
    up, xin = ops.upsample_noise_pack(prev, size, amp=0.2, seed=1234, sample_base=3)
    ref = net._run_block(net.body[3], xin, up, "t", None).numpy().copy()
    blk = fused.desc.body[3]
    nb = int(hp.lib.hpvg_block_fwd_eval_workspace(ctypes.byref(blk), 2, *size))
    ws = hp.Tensor((nb // 4 + 1,), hp.F32)
    out = hp.Tensor(ref.shape, hp.F32)
    rc = hp.lib.hpvg_block_fwd_eval(ctypes.byref(blk), 3, 2, size[0], size[1], size[2], xin.ptr, 8, up.ptr, out.ptr, ws.ptr,
                                    nb, None)
    assert rc == 0, hp.lib.hpvg_last_error()
    hp.device_sync()
    assert np.array_equal(out.numpy(), ref)
    # too small a workspace is refused before any launch
    assert hp.lib.hpvg_block_fwd_eval(ctypes.byref(blk), 3, 2, size[0], size[1], size[2], xin.ptr, 8, up.ptr, out.ptr,
                                      ws.ptr, 16, None) != 0


def test_fused_entry_replayed_as_a_cuda_graph(hpvg_gpu):
    """FusedSampler(graph=True): one capture per batch size, then every batch is a D2D of z, a counter update and a
    graph launch — the same bits as the direct call, for changing z and sample indices (forwards and backwards)."""
    hp = hpvg_gpu
    from hpvg import sampling
    net, opt = _build(hp, {}, 4, seed=10)
    amps = [1.0, 0.0, 0.0, 0.3, 0.2]
    st = hp.Stream()
    direct = sampling.FusedSampler(net, amps, batch=2, stream=st)
    graphed = sampling.FusedSampler(net, amps, batch=2, stream=st, graph=True)
    rng = np.random.default_rng(12)
    for base in (0, 6, 2, 100):
        z = hp.from_numpy(rng.standard_normal(sampling.z_init_size(opt, 2)).astype(np.float32))
        want = direct(z, sample_base=base, stream=st).numpy().copy()
        got = graphed(z, sample_base=base, stream=st)
        st.sync()
        assert np.array_equal(got.numpy(), want), "sample_base %d" % base
    z1 = hp.from_numpy(rng.standard_normal(sampling.z_init_size(opt, 1)).astype(np.float32))
    want = direct(z1, sample_base=7, stream=st).numpy().copy()          # a second batch size: a second capture
    out = hp.Tensor(want.shape, hp.F32)
    graphed(z1, sample_base=7, out=out, stream=st)
    st.sync()
    assert np.array_equal(out.numpy(), want)
    assert len(graphed._graphs) == 2
    graphed.close()
