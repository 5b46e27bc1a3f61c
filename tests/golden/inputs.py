"""Deterministic synthetic inputs shared by oracle/gen_golden.py (which feeds them to the reference's own code)
and by the tests (which feed them to the oracle restatement and to the CUDA path). Everything is drawn from
numpy's legacy RandomState, whose streams are frozen across numpy versions, so only OUTPUTS need committing.

Box recipe follows the reference's `_demo_mm_inputs` (tests/test_models/test_forward.py:369-444):
cx, cy, bw, bh ~ U(0,1) scaled to the image and clipped.
"""
import numpy as np
import torch

STRIDES = (8, 16, 32, 64, 128)
REGRESS_RANGES = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, 1e8))


def level_sizes(H, W, strides=STRIDES):
    """FPN map sizes of a padded H x W image (conv arithmetic of ResNet + FPN P6/P7 stride-2 3x3 pad-1 convs)."""
    sizes = []
    h, w = H // 8, W // 8
    for i in range(5):
        sizes.append((h, w))
        h, w = (h + 1) // 2, (w + 1) // 2
    return sizes


def demo_boxes(rng, n, H, W):
    cx, cy, bw, bh = rng.rand(n, 4).T
    x1 = ((cx * W) - (W * bw / 2)).clip(0, W)
    y1 = ((cy * H) - (H * bh / 2)).clip(0, H)
    x2 = ((cx * W) + (W * bw / 2)).clip(0, W)
    y2 = ((cy * H) + (H * bh / 2)).clip(0, H)
    return np.stack([x1, y1, x2, y2], 1).astype(np.float32)


def make_gt(seed, B, H, W, max_gt=9, max_ignore=3, num_classes=80, with_ignore=True, empty_first=False,
            duplicate_boxes=False):
    """Per-image GT boxes/labels and ignore boxes."""
    rng = np.random.RandomState(seed)
    gts, labels, ignores = [], [], []
    for b in range(B):
        n = 0 if (empty_first and b == 0) else int(rng.randint(1, max_gt + 1))
        boxes = demo_boxes(rng, n, H, W)
        if duplicate_boxes and n >= 2:
            # two GTs of identical area covering the same points: exercises the first-index tie-break
            boxes[1] = boxes[0]
        gts.append(torch.from_numpy(boxes.reshape(-1, 4)))
        labels.append(torch.from_numpy(rng.randint(0, num_classes, size=n).astype(np.int64)))
        ni = int(rng.randint(0, max_ignore + 1)) if with_ignore else 0
        ignores.append(torch.from_numpy(demo_boxes(rng, ni, H, W).reshape(-1, 4)))
    return gts, labels, (ignores if with_ignore else None)


def make_head_outputs(seed, B, H, W, num_classes=80, train=True, cls_mean=-2.0):
    """Random head outputs (NCHW lists): cls logits ~ N(-2, 1.5), bbox >= 0 (post-ReLU), centerness logits."""
    rng = np.random.RandomState(seed)
    cls, box, ctr = [], [], []
    for lvl, (h, w) in enumerate(level_sizes(H, W)):
        cls.append(torch.from_numpy((rng.randn(B, num_classes, h, w) * 1.5 + cls_mean).astype(np.float32)))
        b = np.maximum(rng.randn(B, 4, h, w) * 2.0 + 2.0, 0).astype(np.float32)
        if not train:
            b = b * STRIDES[lvl]
        box.append(torch.from_numpy(b))
        ctr.append(torch.from_numpy(rng.randn(B, 1, h, w).astype(np.float32)))
    return cls, box, ctr


def make_tensor(rng, *shape, scale=1.0):
    return torch.from_numpy((rng.randn(*shape) * scale).astype(np.float32))


def fill_state_dict_(sd, seed, skip=()):
    """Overwrite every tensor of a state_dict, in key order, with seeded values sized by fan-in.
    BatchNorm running_var / weights get positive values; num_batches_tracked is left alone."""
    rng = np.random.RandomState(seed)
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k in skip:
            continue
        if k.endswith("running_var"):
            v.copy_(torch.from_numpy((rng.rand(*v.shape) + 0.5).astype(np.float32)))
        elif k.endswith("running_mean"):
            v.copy_(make_tensor(rng, *v.shape, scale=0.1))
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(make_tensor(rng, *v.shape, scale=(2.0 / fan_in) ** 0.5))
        elif k.endswith(".weight"):  # norm weight
            v.copy_(torch.from_numpy((rng.rand(*v.shape) + 0.5).astype(np.float32)))
        elif k.endswith("scale"):
            v.copy_(torch.tensor(float(rng.rand() + 0.5)))
        else:  # biases
            v.copy_(make_tensor(rng, *v.shape, scale=0.1))
    return sd


def rla_state_dict(seed, layers=(3, 4, 6, 3), rla_channel=32):
    """Seeded state_dict of RLA_ResNet (reference names: mmdet/models/backbones/resnet_rla.py:205-268). Every tensor is
    drawn from its own RandomState(crc32(name) ^ seed), so the values do not depend on key order. Gains are chosen so
    that activations stay O(1) through 16 residual blocks and the tanh of the state update is not saturated."""
    import zlib
    from collections import OrderedDict
    sd = OrderedDict()

    def rs(name):
        return np.random.RandomState((zlib.crc32(name.encode()) ^ seed) & 0x7fffffff)

    def conv(name, o, i, k, gain=1.0):
        sd[name + ".weight"] = make_tensor(rs(name), o, i, k, k, scale=gain * (2.0 / (i * k * k)) ** 0.5)

    def bn(name, c, gain=1.0):
        r = rs(name)
        sd[name + ".weight"] = torch.from_numpy(((r.rand(c) + 0.5) * gain).astype(np.float32))
        sd[name + ".bias"] = make_tensor(r, c, scale=0.1)
        sd[name + ".running_mean"] = make_tensor(r, c, scale=0.1)
        sd[name + ".running_var"] = torch.from_numpy((r.rand(c) + 0.5).astype(np.float32))

    conv("conv1", 64, 3, 7)
    bn("bn1", 64)
    inpl = 64
    for li, nb in enumerate(layers):
        planes = 64 * 2 ** li
        conv(f"conv_outs.{li}", rla_channel, planes * 4, 1, gain=0.5)
        conv(f"recurrent_convs.{li}", rla_channel, rla_channel, 3)
        for bi in range(nb):
            p = f"stages.{li}.{bi}"
            conv(p + ".conv1", planes, inpl + rla_channel, 1)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4, gain=0.3)
            if bi == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1, gain=0.7)
                bn(p + ".downsample.1", planes * 4)
            bn(f"stage_bns.{li}.{bi}", rla_channel)
            inpl = planes * 4
    return sd


def fill_by_name_(sd, seed, skip_prefix=None):
    """Like fill_state_dict_, but every tensor is drawn from RandomState(crc32(key) ^ seed): independent of key order, so
    a test can rebuild the same values from names and shapes alone."""
    import zlib
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or (skip_prefix and k.startswith(skip_prefix)):
            continue
        rng = np.random.RandomState((zlib.crc32(k.encode()) ^ seed) & 0x7fffffff)
        if k.endswith("running_var"):
            t = torch.from_numpy((rng.rand(*v.shape) + 0.5).astype(np.float32))
        elif k.endswith("running_mean"):
            t = make_tensor(rng, *v.shape, scale=0.1)
        elif v.dim() == 4:
            t = make_tensor(rng, *v.shape, scale=(2.0 / (v.shape[1] * v.shape[2] * v.shape[3])) ** 0.5)
        elif k.endswith(".weight"):
            t = torch.from_numpy((rng.rand(*v.shape) + 0.5).astype(np.float32))
        elif k.endswith("scale"):
            t = torch.tensor(float(rng.rand() + 0.5))
        else:
            t = make_tensor(rng, *v.shape, scale=0.1)
        with torch.no_grad():
            v.copy_(t.reshape(v.shape))
    return sd


def rla_detector_state(seed_bb, seed_rest, num_classes=80):
    """Seeded state_dict (reference names) of FCOS with the RLA_ResNet backbone: rla_state_dict for the backbone,
    fill_by_name_ for FPN + head (names / shapes from dsl_b200.params), classification prior bias -4.59 and a small
    conv_cls gain so that the focal loss is in its working range."""
    from collections import OrderedDict
    from dsl_b200.params import fpn_spec, head_spec
    sd = OrderedDict(("backbone." + k, v) for k, v in rla_state_dict(seed_bb).items())
    rest = OrderedDict((p.name, torch.zeros(p.shape)) for p in fpn_spec() + head_spec(num_classes))
    fill_by_name_(rest, seed_rest)
    rest["bbox_head.conv_cls.weight"] *= 0.1
    rest["bbox_head.conv_cls.bias"].fill_(-4.59)
    sd.update(rest)
    return sd


# parameters whose gradients tests/golden/rla_detector.npz samples (oracle/gen_golden.py::gen_rla_detector)
RLA_DET_GRAD_KEYS = ("backbone.stages.1.0.bn1.weight", "backbone.conv_outs.3.weight", "backbone.stage_bns.2.1.bias",
                     "backbone.stages.2.2.conv1.weight", "neck.lateral_convs.0.conv.weight", "bbox_head.conv_cls.bias",
                     "bbox_head.reg_convs.1.gn.weight", "bbox_head.scales.2.scale")
