"""Host-side mirror of the reference's `src/modules/networks_2d.py` (the image path).

Same class names / constructor arguments / `construct` signatures / parameter names as the reference
(ConvBlock2D :44, ConvBlock2DSN :56, FeatureExtractor :75, Encode2DVAE :85, WDiscriminator2D :162,
GeneratorHPVAEGAN :188).  There is no separate 2-D kernel: an (N,C,H,W) tensor is the T == 1 case of the 3-D
channels-last layout, a (Cout,Cin,3,3) filter is packed into the centre temporal tap of the tcgen05 filter bank
(`hpvg_conv_pack_weights(kt=1)`), `ops.ResizeBilinear(size, align_corners=True)` (src/utils/images.py:40-51) is the
T == 1 case of the linear-resize kernels.  Differences from the 3-D graph that ARE mirrored:
  * BatchNorm2d parameters are named `...1.gamma` (no `bn2d.` level, src/tools/pt2ms.py:74);
  * refinement noise is added at EVERY scale in random mode — there is no `vae_levels` gate
    (networks_2d.py:274-277 vs networks_3d.py:443-446);
  * pyramid geometry is `[int(s*ar), s]` (images.py:110-119)."""
from . import networks_3d as n3
from .networks_3d import Cell, ConvLayer, Sequential, as4d, as5d  # noqa: F401
from .utils import images as uimg


def ConvBlock2D(in_channel, out_channel, ker_size=3, padding=1, stride=1, bn=True, act="lrelu", rng=None):
    """networks_2d.py:44-53."""
    return n3.ConvBlock3D(in_channel, out_channel, ker_size, padding, stride, bn=bn, act=act, rng=rng, kt=1,
                          bn_prefix="1.")


def ConvBlock2DSN(in_channel, out_channel, ker_size=3, padding=1, stride=1, bn=True, act="lrelu", rng=None):
    """networks_2d.py:56-72."""
    return n3.ConvBlock3DSN(in_channel, out_channel, ker_size, padding, stride, bn=bn, act=act, rng=rng, kt=1)


class FeatureExtractor(n3.FeatureExtractor):
    """networks_2d.py:75-82."""

    def __init__(self, in_channel, out_channel, ker_size, padding, stride, num_blocks=2, return_linear=False, rng=None):
        super().__init__(in_channel, out_channel, ker_size, padding, stride, num_blocks, return_linear, rng, kt=1)


class Encode2DVAE(n3.Encode3DVAE):
    """networks_2d.py:85-107."""
    KT = 1


class WDiscriminator2D(n3.WDiscriminator3D):
    """networks_2d.py:162-185."""
    KT = 1


class GeneratorHPVAEGAN(n3.GeneratorHPVAEGAN):
    """networks_2d.py:188-282.  `construct` takes / returns (N,C,H,W) tensors."""
    KT = 1
    BN_PREFIX = "1."
    ENCODER = Encode2DVAE

    def stage_shape(self, index):
        h, w = uimg.scale_shape_2d(self.opt, index)
        return (1, h, w)

    def noise_at(self, index, is_random):
        return bool(is_random)      # networks_2d.py:274-277: every scale in random mode
