"""Host-side logic of the multi-GPU sampling path on CPU: sample sharding, the equal-count padding the
all-gather needs, un-sharding into global sample order, a real world_size-2 gloo run of that exchange, the Fréchet
distance, the checkpoint key map, and the rule that the product package never imports torch or the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("num,batch,world", [(512, 16, 8), (512, 16, 1), (10, 4, 2), (7, 3, 4), (1, 8, 2), (33, 8, 3)])
def test_shard_unshard_roundtrip(num, batch, world):
    from hpvg import dist
    per_rank = [dist.shard_rows(num, batch, r, world) for r in range(world)]
    assert len({len(p) for p in per_rank}) == 1                       # equal counts on every rank
    owned = sorted(i for p in per_rank for i in p if i >= 0)
    assert owned == list(range(num))                                  # a partition: every sample exactly once
    rows = np.concatenate([np.array([[i, 2 * i + 1] if i >= 0 else [-9, -9] for i in p], np.float32) for p in per_rank])
    out = dist.unshard_rows(rows, num, batch, world)
    assert np.array_equal(out[:, 0], np.arange(num)) and np.array_equal(out[:, 1], 2 * np.arange(num) + 1)


def test_sample_noise_is_keyed_by_global_index_not_by_rank():
    from hpvg import sampling
    a = sampling.host_noise_for_sample(3, 17, (4, 5))
    assert np.array_equal(a, sampling.host_noise_for_sample(3, 17, (4, 5)))
    assert not np.array_equal(a, sampling.host_noise_for_sample(3, 18, (4, 5)))
    # any partition generates the same set of (index -> noise): chunks only group indices
    for world in (1, 2, 8):
        idx = sorted(i for r in range(world) for c in sampling.local_chunks(37, 4, r, world) for i in c)
        assert idx == list(range(37))


def _gloo_worker(rank, world, port, num, batch, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dist_util import GlooCommunicator
    from hpvg import dist as hd
    comm = GlooCommunicator()
    idx = hd.shard_rows(num, batch, comm.rank, comm.world)
    rows = np.array([[i, i * i, 7.0] if i >= 0 else [0, 0, 0] for i in idx], np.float32)   # "moments" of sample i
    gathered = comm.all_gather_rows(rows)
    out = hd.unshard_rows(gathered, num, batch, comm.world)
    comm.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_single_process():
    import torch.multiprocessing as mp
    num, batch, world = 21, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, num, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[i, i * i, 7.0] for i in range(num)], np.float32)
    for r in range(world):
        assert np.array_equal(res[r], want)          # identical, complete and in global order on every rank


def test_frechet_distance_known_answers():
    from hpvg import fid
    rng = np.random.default_rng(0)
    a = rng.standard_normal((500, 8))
    mu, sig = a.mean(0), np.cov(a, rowvar=False)
    assert abs(fid.calculate_frechet_distance(mu, sig, mu, sig)) < 1e-6
    # commuting (diagonal) covariances: closed form sum (sqrt(a)-sqrt(b))^2 + |dmu|^2
    d1, d2 = np.array([1.0, 4.0, 9.0]), np.array([4.0, 1.0, 16.0])
    m1, m2 = np.zeros(3), np.array([1.0, 2.0, 2.0])
    want = float(((np.sqrt(d1) - np.sqrt(d2)) ** 2).sum() + 9.0)
    assert abs(fid.calculate_frechet_distance(m1, np.diag(d1), m2, np.diag(d2)) - want) < 1e-8
    # moments_to_stats == np.mean / np.cov(rowvar=False) (fid_score.py:176-177)
    f = rng.standard_normal((300, 64))
    row = np.concatenate([f.sum(0), (f.T @ f).ravel()])
    mu2, sig2 = fid.moments_to_stats(row, 300)
    assert np.allclose(mu2, f.mean(0)) and np.allclose(sig2, np.cov(f, rowvar=False))


def test_pytorch_key_map_and_reference_name_translation():
    """src/tools/pt2ms.py:129-188: our map lands on THIS PACKAGE's positional names; the reference's MindSpore names differ
    for generator stages >= 1 (`body.0.0.N.`, pt2ms.py:156-157) and discriminator blocks >= 1 (`body.0.j.`, :112-113) and
    are translated both ways."""
    from hpvg import checkpoint as ck
    z = np.zeros
    state = {"encode.features.conv_block_0.conv.weight_orig": z((64, 3, 3, 3, 3)),
             "encode.features.conv_block_0.conv.weight_u": z((64,)),
             "encode.features.conv_block_2.conv.bias": z((64,)),
             "encode.mu.conv.weight": z((128, 64, 3, 3, 3)),
             "encode.logvar.conv.bias": z((128,)),
             "decoder.head.conv.weight": z((64, 128, 3, 3, 3)),
             "decoder.head.norm.running_mean": z((64,)),
             "decoder.head.norm.num_batches_tracked": z(()),
             "decoder.block3.norm.weight": z((64,)),
             "decoder.tail.weight": z((3, 64, 3, 3, 3)),
             "body.2.block0.conv.bias": z((64,)),
             "body.2.head.norm.running_var": z((64,)),
             "body.2.tail.bias": z((3,))}
    got = ck.p2m_HPVAEGAN_3d({"state_dict": state})
    assert set(got) == {"encode._features.0.0.weight", "encode._features.0.0.weight_u", "encode._features.2.0.bias",
                        "encode._mu.0.weight", "encode._logvar.0.bias", "decoder.0.0.weight",
                        "decoder.0.1.bn2d.moving_mean", "decoder.4.1.bn2d.gamma", "decoder.6.weight",
                        "body.2.1.0.bias", "body.2.0.1.bn2d.moving_variance", "body.2.6.bias"}
    assert got["encode._features.0.0.weight_u"].shape == (64, 1)
    got2 = ck.p2m_HPVAEGAN_2d(state)
    assert "decoder.0.1.moving_mean" in got2 and "decoder.4.1.gamma" in got2
    # the names pt2ms.py itself produces for the same PyTorch keys (its string substitutions restated literally)
    ref_names = ck.to_reference_names(got)
    assert "body.0.0.2.1.0.bias" in ref_names and "body.0.0.2.0.1.bn2d.moving_variance" in ref_names \
        and "body.0.0.2.6.bias" in ref_names and "decoder.0.0.weight" in ref_names
    assert ck.from_reference_names(ref_names).keys() == got.keys()
    assert ck.from_reference_names(got).keys() == got.keys()              # idempotent on our names
    # stage 0 keys are identical in both namings and must not be mistaken for a later stage
    s0 = {"body.0.0.0.weight": 1, "body.0.0.1.bn2d.gamma": 2, "body.0.6.weight": 3, "body.0.0.1.0.0.weight": 4}
    assert ck.from_reference_names(s0) == {"body.0.0.0.weight": 1, "body.0.0.1.bn2d.gamma": 2, "body.0.6.weight": 3,
                                           "body.1.0.0.weight": 4}
    # discriminator (pt2ms.py:105-126)
    dstate = {"head.conv.weight_orig": z((64, 3, 3, 3, 3)), "head.conv.weight_u": z((64,)), "body.block0.conv.bias": z((64,)),
              "body.block3.conv.weight_orig": z((64, 64, 3, 3, 3)), "tail.weight": z((1, 64, 3, 3, 3)), "tail.bias": z((1,))}
    dgot = ck.p2m_WDiscriminator_3d({"state_dict": dstate})
    assert set(dgot) == {"head.0.weight", "head.0.weight_u", "body.0.0.bias", "body.3.0.weight", "tail.weight", "tail.bias"}
    dref = ck.to_reference_names(dgot)
    assert set(dref) == {"head.0.weight", "head.0.weight_u", "body.0.0.bias", "body.0.3.0.weight", "tail.weight", "tail.bias"}
    assert set(ck.from_reference_names(dref)) == set(dgot)


def test_product_package_does_not_import_torch_or_oracle():
    code = ("import sys; sys.path.insert(0, %r); import hpvg, hpvg.networks_3d, hpvg.networks_2d, hpvg.train, "
            "hpvg.sampling, hpvg.dist, hpvg.fid, hpvg.driver, hpvg.checkpoint; "
            "bad = [m for m in ('torch', 'oracle', 'triton') if m in sys.modules]; assert not bad, bad"
            % os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
    subprocess.check_call([sys.executable, "-c", code])


def test_mindspore_ckpt_container_reader_and_writer(tmp_path):
    """checkpoint.proto (Checkpoint{repeated Value{tag, TensorProto{dims, tensor_type, tensor_content}}}): the reader
    against an INDEPENDENT hand-rolled encoder of the schema (unpacked and packed dims, a tensor split over two Values,
    an unknown extra field), and the package's own writer round trip — with the reference's parameter names."""
    from hpvg import checkpoint as ck

    def vi(v):
        out = b""
        while True:
            b, v = v & 0x7F, v >> 7
            out += bytes([b | (0x80 if v else 0)])
            if not v:
                return out

    def ld(num, payload):
        return vi((num << 3) | 2) + vi(len(payload)) + payload

    rng = np.random.default_rng(0)
    w = rng.standard_normal((4, 3, 3, 3, 3)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    step = np.array([7], np.int32)
    raw = w.tobytes()
    dims_unpacked = b"".join(vi((1 << 3) | 0) + vi(d) for d in w.shape)
    dims_packed = ld(1, b"".join(vi(d) for d in b.shape))
    blob = b""
    for part in (raw[:200], raw[200:]):          # a tensor split over two Values with the same tag
        blob += ld(1, ld(1, b"body.0.0.1.0.0.weight") + ld(2, dims_unpacked + ld(2, b"Float32") + ld(3, part)))
    blob += ld(1, ld(1, b"body.0.0.1.0.0.bias") + ld(2, dims_packed + ld(2, b"Float32") + ld(3, b.tobytes())))
    blob += ld(1, ld(1, b"global_step") + ld(2, vi((1 << 3) | 0) + vi(1) + ld(2, b"Int32") + ld(3, step.tobytes()))
               + vi((9 << 3) | 0) + vi(5))      # an unknown field is skipped
    path = tmp_path / "netG_1.ckpt"
    path.write_bytes(blob)
    got = ck.load_checkpoint(str(path))
    assert list(got) == ["body.0.0.1.0.0.weight", "body.0.0.1.0.0.bias", "global_step"]
    assert np.array_equal(got["body.0.0.1.0.0.weight"], w) and np.array_equal(got["body.0.0.1.0.0.bias"], b)
    assert got["global_step"].dtype == np.int32 and got["global_step"][0] == 7
    assert set(ck.from_reference_names({k: v for k, v in got.items() if k != "global_step"})) == \
        {"body.1.0.0.weight", "body.1.0.0.bias"}
    out = ck.save_mindspore_ckpt({"a.weight": w, "a.bias": b}, str(tmp_path / "rt.ckpt"), slice_bytes=128)
    back = ck.load_mindspore_ckpt(out)
    assert np.array_equal(back["a.weight"], w) and np.array_equal(back["a.bias"], b)
    (tmp_path / "bad.ckpt").write_bytes(blob[:-3])
    import pytest
    with pytest.raises((ck.HpvgError, IndexError)):
        ck.load_mindspore_ckpt(str(tmp_path / "bad.ckpt"))
