"""Checkpoint / state I/O with the reference's parameter names.

The reference saves one `netG_{i}.ckpt` / `netD_{i}.ckpt` per scale plus `intermediate.json`
(`train_video.py:224-227`, `src/utils/saver.py:55-76`) and can import the original PyTorch `.pth` through the key map of
`src/tools/pt2ms.py:129-188`.  MindSpore's `.ckpt` is a protobuf that only MindSpore can parse, so the container here
is `.npz` — ONE array per parameter.  A maintainer converts a reference checkpoint with
`np.savez(path, **{k: v.asnumpy() for k, v in mindspore.load_checkpoint(f).items()})`.

Parameter names.  This package names parameters by their position in the graph: `body.<stage>.<layer>.<module>.<param>`
for the generator (`body.2.1.0.bias`), `body.<block>.0.<param>` for the discriminator.  The reference's MindSpore
names are the same for `encode.*`, `decoder.*`, `head.*`, `tail.*`, generator stage 0 and discriminator block 0, but
DIFFER for every later stage / block because of how its cells get their names (`init_next_stage` deep-copies a stage
whose parameters are already named, networks_3d.py:404; `src/tools/pt2ms.py:156-157, 112-113` documents the result):
    generator stage N >= 1        reference `body.0.0.N.<layer>...`      here `body.N.<layer>...`
    discriminator block j >= 1    reference `body.0.j.0.<param>`         here `body.j.0.<param>`
`from_reference_names` / `to_reference_names` translate between the two; `load_param_into_net` accepts either and
`save_checkpoint(..., reference_names=True)` writes the reference's."""
import json
import os
import re

import numpy as np

from .runtime import HpvgError


def state_dict(cell, stream=None):
    """name -> float32 numpy array (device -> host copy of every parameter of `cell`).  stream=None waits for every
    stream of the device first (Tensor.numpy), so weights being updated on a caller's stream are read after the update."""
    return {k: t.numpy(stream) for k, t in cell.parameters_dict().items()}


_G_STAGE = re.compile(r"^body\.0\.0\.(\d+)\.(\d+)\.(.+)$")     # reference: generator stage N >= 1
_D_BLOCK = re.compile(r"^body\.0\.(\d+)\.0\.([a-z_]+)$")       # reference: discriminator block j >= 1
_OUR_G = re.compile(r"^body\.(\d+)\.(\d+)\.(.+)$")
_OUR_D = re.compile(r"^body\.(\d+)\.0\.([a-z_]+)$")


def is_discriminator_state(params):
    return any(k.startswith("head.") or k.startswith("tail.") for k in params)


def from_reference_names(params):
    """Reference (MindSpore) parameter names -> this package's (see the module docstring).  Idempotent on our names."""
    disc = is_discriminator_state(params)
    out = {}
    for k, v in params.items():
        new = k
        if disc:
            m = _D_BLOCK.match(k)
            if m and int(m.group(1)) >= 1:
                new = "body.%s.0.%s" % (m.group(1), m.group(2))
        else:
            m = _G_STAGE.match(k)
            if m and int(m.group(1)) >= 1:
                new = "body.%s.%s.%s" % m.groups()
        if new in out:
            raise HpvgError("parameter names collide after translation: %s" % new)
        out[new] = v
    return out


def to_reference_names(params):
    """This package's parameter names -> the reference's MindSpore names (inverse of from_reference_names)."""
    disc = is_discriminator_state(params)
    out = {}
    for k, v in params.items():
        new = k
        if disc:
            m = _OUR_D.match(k)
            if m and int(m.group(1)) >= 1:
                new = "body.0.%s.0.%s" % (m.group(1), m.group(2))
        else:
            m = _OUR_G.match(k)
            if m and int(m.group(1)) >= 1 and not _G_STAGE.match(k):
                new = "body.0.0.%s.%s.%s" % m.groups()
        out[new] = v
    return out


def save_checkpoint(cell, filename, reference_names=False, stream=None):
    """saver.py:55-57.  Writes `<filename>` (suffix forced to .npz); reference_names=True stores the arrays under the
    reference's MindSpore parameter names."""
    filename = _npz(filename)
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    sd = state_dict(cell, stream)
    np.savez(filename, **(to_reference_names(sd) if reference_names else sd))
    return filename


def load_checkpoint(filename):
    """saver.py:59-63 -> {name: array}."""
    with np.load(_npz(filename)) as f:
        return {k: f[k] for k in f.files}


def load_param_into_net(cell, params, strict=True):
    """mindspore.load_param_into_net: match by name; returns the list of parameters that were NOT loaded."""
    mine = cell.parameters_dict()
    params = from_reference_names(params)      # a converted reference .ckpt loads as it is
    missing = [k for k in mine if k not in params]
    if strict and missing:
        raise HpvgError("checkpoint is missing %d parameters, e.g. %s" % (len(missing), missing[:3]))
    cell.load_parameters(params, strict=False)
    return missing


def save_json(obj, filename):
    """saver.py:65-68 (`intermediate.json`: {'noise_amps': [...], 'scale_idx': i}, train_video.py:224)."""
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    with open(filename, "w+") as f:
        json.dump(obj, f)


def load_json(filename):
    with open(filename, "r") as f:
        return json.load(f)


def _npz(filename):
    root, ext = os.path.splitext(filename)
    return filename if ext == ".npz" else root + ".npz"


# ---------------------------------------------------------------------------------------------------------------------
# Original-PyTorch key map (pt2ms.py:129-188 for 3-D, :30-89 for 2-D): `state` maps the PyTorch HP-VAE-GAN names
# (encode.features.conv_block_i.conv.weight_orig, decoder.head.conv.weight, body.s.blockj.norm.running_mean, ...) to
# arrays; the result uses this package's positional names (to_reference_names() gives the MindSpore reference's).
# ---------------------------------------------------------------------------------------------------------------------
_BN = {"weight": "gamma", "bias": "beta", "running_mean": "moving_mean", "running_var": "moving_variance"}


def _p2m(state, bn_prefix, n_layers):
    out = {}
    for key, value in state.items():
        if key.endswith("num_batches_tracked"):
            continue
        parts = key.split(".")
        new = []
        i = 0
        top = parts[0]
        if top == "encode":
            new.append("encode")
            i = 1
            if parts[i] == "features":
                m = re.fullmatch(r"conv_block_(\d+)", parts[i + 1])
                if not m:
                    raise HpvgError("unexpected encoder key %s" % key)
                new += ["_features", m.group(1)]
                i += 2
            elif parts[i] in ("mu", "logvar"):
                new.append("_" + parts[i])
                i += 1
            rest = parts[i:]
            if rest[0] == "conv":
                rest = ["0"] + rest[1:]
            rest = ["weight" if r == "weight_orig" else r for r in rest]
            new += rest
        elif top in ("decoder", "body"):
            new.append(top)
            i = 1
            if top == "body":
                new.append(parts[i])      # stage index
                i += 1
            blk = parts[i]
            if blk == "head":
                new.append("0")
            elif blk == "tail":
                new.append(str(n_layers + 1))
            else:
                m = re.fullmatch(r"block(\d+)", blk)
                if not m:
                    raise HpvgError("unexpected block key %s" % key)
                new.append(str(int(m.group(1)) + 1))
            rest = parts[i + 1:]
            if rest and rest[0] == "conv":
                new += ["0"] + rest[1:]
            elif rest and rest[0] == "norm":
                new += [bn_prefix.rstrip(".")] if bn_prefix == "1." else ["1", "bn2d"]
                new.append(_BN.get(rest[1], rest[1]))
            else:
                new += rest              # plain tail conv: decoder.tail.weight -> decoder.6.weight
        else:
            new = parts
        arr = np.asarray(value, np.float32)
        if new[-1] in ("weight_u", "weight_v") and arr.ndim == 1:
            arr = arr[:, None]           # pt2ms.py:185-186
        out[".".join(new)] = arr
    return out


def p2m_HPVAEGAN_3d(state, num_layer=5):
    """pt2ms.py:129-188."""
    return _p2m(state.get("state_dict", state), "1.bn2d.", num_layer)


def p2m_HPVAEGAN_2d(state, num_layer=5):
    """pt2ms.py:30-89."""
    return _p2m(state.get("state_dict", state), "1.", num_layer)


def p2m_WDiscriminator(state):
    """pt2ms.py:8-28 (2-D) / :105-126 (3-D): PyTorch WDiscriminator names (head.conv.weight_orig, body.block2.conv.bias,
    tail.weight, ...) -> this package's names (head.0.weight, body.2.0.bias, tail.weight)."""
    out = {}
    for key, value in state.get("state_dict", state).items():
        parts = key.split(".")
        new = []
        for q in parts:
            m = re.fullmatch(r"block(\d+)", q)
            if m:
                new.append(m.group(1))
            elif q == "conv":
                new.append("0")
            elif q == "weight_orig":
                new.append("weight")
            else:
                new.append(q)
        arr = np.asarray(value, np.float32)
        if new[-1] in ("weight_u", "weight_v") and arr.ndim == 1:
            arr = arr[:, None]
        out[".".join(new)] = arr
    return out


p2m_WDiscriminator_3d = p2m_WDiscriminator
p2m_WDiscriminator_2d = p2m_WDiscriminator
