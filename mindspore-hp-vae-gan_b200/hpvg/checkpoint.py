"""Checkpoint / state I/O with the reference's parameter names.

The reference saves one `netG_{i}.ckpt` / `netD_{i}.ckpt` per scale plus `intermediate.json`
(`train_video.py:224-227`, `src/utils/saver.py:55-76`) and can import the original PyTorch `.pth` through the key map of
`src/tools/pt2ms.py:129-188`.  The native container here is `.npz` — ONE array per parameter; MindSpore's `.ckpt` (a small
protobuf schema) is read and written directly by `load_mindspore_ckpt` / `save_mindspore_ckpt` at the bottom of this file,
so a reference checkpoint loads as it is: `load_param_into_net(netG, load_checkpoint("netG_9.ckpt"))`.

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
    """saver.py:59-63 -> {name: array}.  A file ending in .ckpt that exists is read as a MindSpore checkpoint
    (load_mindspore_ckpt); everything else is this package's .npz container."""
    if filename.endswith(".ckpt") and os.path.exists(filename):
        return load_mindspore_ckpt(filename)
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


# ---------------------------------------------------------------------------------------------------------------------
# MindSpore `.ckpt` container without MindSpore (eval_video.py:162-170 `mindspore.load_checkpoint`).
# A .ckpt is a protobuf (mindspore/ccsrc/utils/checkpoint.proto, proto2, unchanged across 1.x / 2.x):
#     message Checkpoint  { repeated Value value = 1; }
#     message Value       { required string tag = 1; required TensorProto tensor = 2; }
#     message TensorProto { repeated int64 dims = 1; required string tensor_type = 2; required bytes tensor_content = 3; }
# `tensor_type` is the MindSpore dtype name ("Float32", "Float16", "Int32", ...).  mindspore.save_checkpoint splits a
# large tensor over several consecutive Values with the same tag (the dims are those of the whole tensor); their
# contents are concatenated here.  The wire format is restated by hand (varints, length-delimited fields) — MindSpore is
# not installable offline and the reference ships no .ckpt, so this reader is checked against an independent encoder of
# the same schema (tests/test_cpu_dist.py), not against a file written by MindSpore ("parity unpinned").
# ---------------------------------------------------------------------------------------------------------------------
_MS_DTYPES = {"Float32": np.float32, "Float16": np.float16, "Float64": np.float64, "Int32": np.int32, "Int64": np.int64,
              "Int8": np.int8, "UInt8": np.uint8, "Bool": np.bool_}


def _varint(buf, pos):
    val, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise HpvgError("corrupt varint in checkpoint")


def _fields(buf, pos, end):
    """Yield (field number, wire type, value) of one protobuf message; value = int (varint / fixed) or a memoryview."""
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 1:
            v = int.from_bytes(buf[pos:pos + 8], "little")
            pos += 8
        elif wt == 5:
            v = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        else:
            raise HpvgError("unsupported protobuf wire type %d in checkpoint" % wt)
        if pos > end:
            raise HpvgError("truncated checkpoint")
        yield num, wt, v


def load_mindspore_ckpt(filename):
    """A MindSpore .ckpt -> {parameter name: numpy array} (the reference's names; see from_reference_names)."""
    with open(filename, "rb") as f:
        buf = memoryview(f.read())
    chunks, dims_of, type_of, order = {}, {}, {}, []
    for num, wt, value in _fields(buf, 0, len(buf)):
        if num != 1 or wt != 2:
            continue
        tag, tensor = None, None
        for n2, w2, v2 in _fields(value, 0, len(value)):
            if n2 == 1 and w2 == 2:
                tag = bytes(v2).decode("utf-8")
            elif n2 == 2 and w2 == 2:
                tensor = v2
        if tag is None or tensor is None:
            raise HpvgError("checkpoint Value without tag / tensor")
        dims, ttype, content = [], None, b""
        for n3, w3, v3 in _fields(tensor, 0, len(tensor)):
            if n3 == 1 and w3 == 0:
                dims.append(v3 if v3 < (1 << 63) else v3 - (1 << 64))
            elif n3 == 1 and w3 == 2:          # packed encoding of the repeated field
                p = 0
                while p < len(v3):
                    d, p = _varint(v3, p)
                    dims.append(d)
            elif n3 == 2 and w3 == 2:
                ttype = bytes(v3).decode("utf-8")
            elif n3 == 3 and w3 == 2:
                content = v3
        if tag not in chunks:
            chunks[tag], dims_of[tag], type_of[tag] = [], dims, ttype
            order.append(tag)
        chunks[tag].append(bytes(content))
    out = {}
    for tag in order:
        dt = _MS_DTYPES.get(type_of[tag])
        if dt is None:
            raise HpvgError("checkpoint tensor %s has unsupported dtype %r" % (tag, type_of[tag]))
        arr = np.frombuffer(b"".join(chunks[tag]), dtype=dt)
        shape = tuple(int(d) for d in dims_of[tag])
        if int(np.prod(shape, dtype=np.int64)) != arr.size:
            raise HpvgError("checkpoint tensor %s: %d elements for dims %s" % (tag, arr.size, shape))
        out[tag] = arr.reshape(shape).copy()
    return out


def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def save_mindspore_ckpt(params, filename, slice_bytes=None):
    """{name: array} -> a file in MindSpore's .ckpt wire format (float32 tensors), so weights trained here can be handed
    back to the reference's `mindspore.load_checkpoint`.  slice_bytes: split tensor contents like save_checkpoint does."""
    names = {np.dtype(v): k for k, v in _MS_DTYPES.items()}
    with open(filename, "wb") as f:
        for tag, arr in params.items():
            a = np.ascontiguousarray(arr)
            raw = a.tobytes()
            step = slice_bytes or max(len(raw), 1)
            for off in range(0, max(len(raw), 1), step):
                tensor = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in a.shape)
                tensor += _enc_field(2, names[a.dtype].encode()) + _enc_field(3, raw[off:off + step])
                f.write(_enc_field(1, _enc_field(1, tag.encode("utf-8")) + _enc_field(2, tensor)))
    return filename
