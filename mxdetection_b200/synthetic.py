"""Synthetic COCO-shaped inputs of the BASELINE.json configs.

NumPy only (no kernels, no oracle): the single definition of the workloads
used by tests/ and bench.py.  All draws come from
``np.random.default_rng(1000*cfg + image_index)``.
"""
import numpy as np

F = np.float32
FPN_STRIDES = (4, 8, 16, 32, 64)


def fpn_shapes(img_h, img_w, strides=FPN_STRIDES):
    """ceil(img/stride) per level: 800x1088 -> (200,272),(100,136),(50,68),(25,34),(13,17)."""
    return [(-(-img_h // s), -(-img_w // s)) for s in strides]


def gt_boxes(rng, img_h, img_w, num=None):
    g = int(rng.integers(1, 101)) if num is None else int(num)
    side = np.exp(rng.uniform(np.log(16), np.log(600), g))
    asp = np.exp(rng.uniform(np.log(0.33), np.log(3), g))
    w = side * np.sqrt(asp); h = side / np.sqrt(asp)
    cx = rng.uniform(0, img_w, g); cy = rng.uniform(0, img_h, g)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
    return b.astype(F)


def rois_for_image(rng, img_h, img_w, num, gts, batch_index=0):
    """25 % jittered GTs (+-10 % size, +-10 % shift), 75 % log-uniform scale 16-800 px, aspect .5-2, clipped."""
    n_gt = num // 4
    src = gts[rng.integers(0, len(gts), n_gt)]
    w = src[:, 2] - src[:, 0]; h = src[:, 3] - src[:, 1]
    cx = (src[:, 0] + src[:, 2]) / 2 + rng.uniform(-0.1, 0.1, n_gt) * w
    cy = (src[:, 1] + src[:, 3]) / 2 + rng.uniform(-0.1, 0.1, n_gt) * h
    w = w * rng.uniform(0.9, 1.1, n_gt); h = h * rng.uniform(0.9, 1.1, n_gt)
    a = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    m = num - n_gt
    s = np.exp(rng.uniform(np.log(16), np.log(800), m)); asp = np.exp(rng.uniform(np.log(0.5), np.log(2), m))
    w = s * np.sqrt(asp); h = s / np.sqrt(asp)
    cx = rng.uniform(0, img_w, m); cy = rng.uniform(0, img_h, m)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    r = np.concatenate([a, b], 0)
    r[:, 0::2] = np.clip(r[:, 0::2], 0, img_w - 1); r[:, 1::2] = np.clip(r[:, 1::2], 0, img_h - 1)
    out = np.empty((num, 5), F)
    out[:, 0] = batch_index
    out[:, 1:] = r.astype(F)
    return out[rng.permutation(num)]


def cfg1(seed=1000):
    """RoIAlign 7x7, 512 RoIs on one 256x200x272 map (800x1088 image, stride 4)."""
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((1, 256, 200, 272), dtype=F)
    gts = gt_boxes(rng, 800, 1088)
    rois = rois_for_image(rng, 800, 1088, 512, gts, 0)
    grad_out = rng.standard_normal((512, 256, 7, 7), dtype=F)
    return dict(data=data, rois=rois, grad_out=grad_out, pooled=(7, 7), scale=0.25, sample_ratio=2)


def rpn_inputs(cfg, batch, img_h, img_w, num_anchors=3, strides=FPN_STRIDES, first_image=0):
    """Per level: scores (B, H*W*A) = sigmoid(N(-4,2)), deltas (B, H*W*A, 4) = N(0,0.25); (y,x,a) order."""
    shapes = fpn_shapes(img_h, img_w, strides)
    scores = [np.empty((batch, h * w * num_anchors), F) for h, w in shapes]
    deltas = [np.empty((batch, h * w * num_anchors, 4), F) for h, w in shapes]
    for b in range(batch):
        rng = np.random.default_rng(1000 * cfg + first_image + b)
        for l, (h, w) in enumerate(shapes):
            n = h * w * num_anchors
            z = rng.normal(-4.0, 2.0, n)
            scores[l][b] = (1.0 / (1.0 + np.exp(-z))).astype(F)
            deltas[l][b] = rng.normal(0.0, 0.25, (n, 4)).astype(F)
    img_shapes = np.tile(np.asarray([[img_h, img_w]], np.int32), (batch, 1))
    return dict(scores=scores, deltas=deltas, feat_shapes=shapes, strides=strides, img_shapes=img_shapes)


def cfg2(batch=2, first_image=0):
    """RPN proposals: 800x1088, 5 levels, A=3 -> 217 413 anchors / image."""
    return rpn_inputs(2, batch, 800, 1088, first_image=first_image)


def fpn_roi_inputs(cfg, batch, img_h, img_w, rois_per_img, channels=256, pooled=(7, 7), first_image=0,
                   with_features=True):
    """FPN RoI stage: 4 maps (strides 4,8,16,32) x `channels`, `rois_per_img` RoIs per image."""
    shapes = fpn_shapes(img_h, img_w, FPN_STRIDES[:4])
    rois, gts_all = [], []
    for b in range(batch):
        rng = np.random.default_rng(1000 * cfg + first_image + b)
        gts = gt_boxes(rng, img_h, img_w, 100 if cfg >= 4 else None)
        gts_all.append(gts)
        rois.append(rois_for_image(rng, img_h, img_w, rois_per_img, gts, b))
    out = dict(rois=np.concatenate(rois, 0), gts=gts_all, feat_shapes=shapes, strides=FPN_STRIDES[:4],
               scales=[1.0 / s for s in FPN_STRIDES[:4]], pooled=pooled, channels=channels, sample_ratio=2)
    if with_features:
        rng = np.random.default_rng(1000 * cfg + 500 + first_image)
        out["feats"] = [rng.standard_normal((batch, channels, h, w), dtype=F) for h, w in shapes]
        out["grad_out"] = rng.standard_normal((batch * rois_per_img, channels) + tuple(pooled), dtype=F)
    return out


def cfg3(batch=8, first_image=0, with_features=True):
    """Faster R-CNN R50-FPN RoI stage: 512 RoIs/img x 8 imgs, 7x7, 1333x800 padded to 800x1344."""
    return fpn_roi_inputs(3, batch, 800, 1344, 512, first_image=first_image, with_features=with_features)


def cfg4_mask(batch=1, first_image=0, with_features=True):
    """Mask branch: 14x14 RoIAlign on 128 RoIs/img, 256 ch."""
    return fpn_roi_inputs(4, batch, 800, 1344, 128, pooled=(14, 14), first_image=first_image,
                          with_features=with_features)


def padded_gts(gts_list, gmax=100):
    B = len(gts_list)
    out = np.zeros((B, gmax, 4), F); num = np.zeros(B, np.int32)
    for b, g in enumerate(gts_list):
        out[b, :len(g)] = g; num[b] = len(g)
    return out, num


def assigner_inputs(cfg, batch, img_h=800, img_w=1344, num_gt=100, first_image=0):
    """cfg4b: all FPN anchors of an 800x1344 image (268 569) vs `num_gt` GTs per image."""
    gts = [gt_boxes(np.random.default_rng(1000 * cfg + 700 + first_image + b), img_h, img_w, num_gt)
           for b in range(batch)]
    g, n = padded_gts(gts, num_gt)
    labels = np.stack([np.random.default_rng(1000 * cfg + 800 + first_image + b).integers(1, 81, num_gt)
                       for b in range(batch)]).astype(np.int32)
    return dict(gts=g, num_gts=n, gt_labels=labels, feat_shapes=fpn_shapes(img_h, img_w), strides=FPN_STRIDES,
                img_shape=(img_h, img_w))
