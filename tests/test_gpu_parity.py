"""GPU parity tests (run on a B200 through gpurun: `pytest -m gpu`).

Every test drives the product path - Python mirror -> ctypes C ABI -> sm_100a kernels - and
compares with the CPU oracle (oracle/*.py NumPy statement, oracle/cpu_ref.c for the large cases)
on identical seeded inputs.  Bars (BASELINE.json north_star / SURVEY.md 8c):
  bit-exact : anchors, flags, IoU, labels, matched-GT indices, top-k indices, NMS keep indices, levels
  RoIAlign  : fwd |d| <= 1e-5*max(1,|ref|), bwd |d| <= 1e-4*max(1,|ref|)
  decode    : <= 1e-5 px (and bit-identical in practice: exp/log are correctly rounded on both sides)
"""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import cref
from oracle.nms import stable_order_desc
from mxdetection_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
F = np.float32
DEV = "cuda"


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().cpu().numpy()


def close(a, ref, tol):
    return np.all(np.abs(a - ref) <= tol * np.maximum(1.0, np.abs(ref)))


@pytest.fixture(scope="module", autouse=True)
def _native_library_is_the_one_running():
    import mxdetection_b200 as m
    before = m.launch_count()
    yield
    assert m.launch_count() > before, "no kernel of libmxdet_sm100.so was launched by the GPU tests"


# =============================================================== RoIAlign (Spec A) ==
@pytest.mark.parametrize("name", ["roi_align_a", "roi_align_b", "roi_align_c"])
def test_roi_align_golden(golden_dir, name, roi_path):
    from mxdetection_b200.ops import roi_align_forward, roi_align_backward
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ps = tuple(int(v) for v in g["pooled"]); sc = float(g["scale"]); sr = int(g["sample_ratio"])
    out = N(roi_align_forward(T(g["data"]), T(g["rois"]), ps, sc, sr))
    assert close(out, g["out"], 1e-5)
    gin = N(roi_align_backward(T(g["grad_out"]), T(g["rois"]), g["data"].shape, ps, sc, sr))
    assert close(gin, g["grad_in"], 1e-4)


@pytest.fixture(params=["ring", "plane", "gather"])
def roi_path(request):
    """The RoIAlign code paths of the library: row-ring forward (forced even for tiny inputs) + tile backward, band
    (plane-resident) forward + tile backward, and gather (workspace NULL)."""
    import mxdetection_b200.ops.roi_align as ra
    old = ra.USE_PLANE_KERNELS
    ra.USE_PLANE_KERNELS = request.param != "gather"
    os.environ.pop("MXD_NO_RING", None); os.environ.pop("MXD_RING_FORCE", None)
    if request.param == "ring":
        os.environ["MXD_RING_FORCE"] = "1"
    elif request.param == "plane":
        os.environ["MXD_NO_RING"] = "1"
    yield request.param
    os.environ.pop("MXD_NO_RING", None); os.environ.pop("MXD_RING_FORCE", None)
    ra.USE_PLANE_KERNELS = old


@pytest.mark.parametrize("sr,ps", [(2, (7, 7)), (2, (14, 14)), (-1, (7, 7)), (4, (7, 7)), (9, (8, 8)), (1, (1, 1))])
def test_roi_align_vs_oracle_random(sr, ps, roi_path):
    from mxdetection_b200.ops import roi_align_forward, roi_align_backward
    rng = np.random.default_rng(7 + sr + ps[0])
    Nn, C, H, W = 2, 37, 50, 68           # C not a multiple of the channel chunk
    data = rng.standard_normal((Nn, C, H, W)).astype(F)
    R = 96
    x1 = rng.uniform(-40, W * 4, R); y1 = rng.uniform(-40, H * 4, R)
    w = np.exp(rng.uniform(np.log(2), np.log(400), R)); h = np.exp(rng.uniform(np.log(2), np.log(400), R))
    rois = np.stack([rng.integers(0, Nn, R), x1, y1, x1 + w, y1 + h], 1).astype(F)
    ref = cref.roi_align_forward(data, rois, ps, 0.25, sr)
    out = N(roi_align_forward(T(data), T(rois), ps, 0.25, sr))
    assert close(out, ref, 1e-5)
    gout = rng.standard_normal(ref.shape).astype(F)
    gref = cref.roi_align_backward(gout, rois, data.shape, ps, 0.25, sr)
    gin = N(roi_align_backward(T(gout), T(rois), data.shape, ps, 0.25, sr))
    assert close(gin, gref, 1e-4)
    # req='add'
    base = rng.standard_normal(data.shape).astype(F)
    acc = T(base.copy())
    roi_align_backward(T(gout), T(rois), data.shape, ps, 0.25, sr, grad_data=acc, accumulate=True)
    assert close(N(acc), base + gref, 1e-4)


@pytest.mark.parametrize("PS", [7, 14])
@pytest.mark.parametrize("C,H,W,R", [(32, 50, 68, 96), (64, 100, 168, 300), (32, 13, 9, 40), (96, 70, 130, 64)])
def test_roi_align_tile_backward_paths(C, H, W, R, PS):
    """Tile-resident backward (C % 32 == 0, 7x7 / 14x14, sr 2): tiny / huge / out-of-image RoIs (those take the
    RED fallback after the tiles are written), partial edge tiles, req=write and req=add."""
    from mxdetection_b200.ops import roi_align_backward
    rng = np.random.default_rng(C + H + R)
    Nn = 2
    x1 = rng.uniform(-30, W * 4, R); y1 = rng.uniform(-30, H * 4, R)
    w = np.exp(rng.uniform(np.log(1), np.log(4 * W), R)); h = np.exp(rng.uniform(np.log(1), np.log(4 * H), R))
    rois = np.stack([rng.integers(0, Nn, R), x1, y1, x1 + w, y1 + h], 1).astype(F)
    rois[0] = [0, 0, 0, 4 * W - 1, 4 * H - 1]          # whole map
    rois[1] = [1, 8, 8, 8.5, 8.5]                        # degenerate: every sample in one pixel quad
    rois[2] = [5, 0, 0, 10, 10]                          # bad batch index: no gradient
    rois[3] = [0, 4 * W - 6, 4 * H - 6, 4 * W - 1, 4 * H - 1]   # clamped at the far border
    gout = rng.standard_normal((R, C, PS, PS)).astype(F)
    shape = (Nn, C, H, W)
    gref = cref.roi_align_backward(gout, rois, shape, (PS, PS), 0.25, 2)
    gin = N(roi_align_backward(T(gout), T(rois), shape, (PS, PS), 0.25, 2))
    assert close(gin, gref, 1e-4)
    base = rng.standard_normal(shape).astype(F)
    acc = T(base.copy())
    roi_align_backward(T(gout), T(rois), shape, (PS, PS), 0.25, 2, grad_data=acc, accumulate=True)
    assert close(N(acc), base + gref, 1e-4)
    # no RoI at all on image 1: its tiles are still zero-filled under req=write
    rois0 = rois.copy(); rois0[:, 0] = 0
    g0 = N(roi_align_backward(T(gout), T(rois0), shape, (PS, PS), 0.25, 2))
    assert np.all(g0[1] == 0) and close(g0, cref.roi_align_backward(gout, rois0, shape, (PS, PS), 0.25, 2), 1e-4)


@pytest.mark.parametrize("seed", range(24))
def test_roi_align_fpn_fuzz_shapes(seed, roi_path):
    """Random FPN geometries through the shipped fast paths (stream forward / tile backward) and their fallbacks:
    odd and 4-aligned widths, 1-4 levels, 32-96 channels, 7x7 and 14x14, RoIs of every size incl. out of image."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    rng = np.random.default_rng(900 + seed)
    L_ = int(rng.integers(1, 5)); Nn = int(rng.integers(1, 4)); C = int(rng.choice([32, 64, 96]))
    PS = int(rng.choice([7, 14]))
    h0 = int(rng.integers(20, 140)); w0 = int(rng.integers(20, 180))
    if seed % 2 == 0:
        w0 = (w0 + 7) // 8 * 8                      # 16-byte aligned rows on the finest levels: TMA band path
    shapes = [(Nn, C, max(2, -(-h0 // (1 << l))), max(2, -(-w0 // (1 << l)))) for l in range(L_)]
    scales = [1.0 / (4 << l) for l in range(L_)]
    feats = [rng.standard_normal(s).astype(F) for s in shapes]
    R = int(rng.integers(1, 200))
    img_h, img_w = 4 * h0, 4 * w0
    x1 = rng.uniform(-20, img_w, R); y1 = rng.uniform(-20, img_h, R)
    bw = np.exp(rng.uniform(np.log(2), np.log(1.2 * img_w), R)); bh = np.exp(rng.uniform(np.log(2), np.log(1.2 * img_h), R))
    rois = np.stack([np.sort(rng.integers(0, Nn, R)), x1, y1, x1 + bw, y1 + bh], 1).astype(F)
    lv = oracle.map_roi_levels(rois, L_) if L_ > 1 else None
    out = roi_align_fpn_forward([T(f) for f in feats], T(rois), (PS, PS), scales, 2)
    ref = cref.roi_align_forward(feats, rois, (PS, PS), scales, 2, lv)
    assert close(N(out), ref, 1e-5)
    gout = rng.standard_normal(ref.shape).astype(F)
    g = roi_align_fpn_backward(T(gout), T(rois), shapes, (PS, PS), scales, 2)
    gref = cref.roi_align_backward(gout, rois, shapes, (PS, PS), scales, 2, lv)
    for a, b in zip(g, gref):
        assert close(N(a), b, 1e-4)


def test_roi_align_edge_cases(roi_path):
    from mxdetection_b200.ops import roi_align_forward, roi_align_backward
    data = np.random.default_rng(0).standard_normal((2, 3, 10, 12)).astype(F)
    rois = np.array([[-1, 0, 0, 10, 10], [0, -500, -500, -400, -400], [1, 44, 36, 47.9, 39.9], [0, 5, 5, 5, 5],
                     [7, 0, 0, 5, 5]], F)
    ref = oracle.roi_align_forward(data, rois, (2, 2), 1.0, 2)
    out = N(roi_align_forward(T(data), T(rois), (2, 2), 1.0, 2))
    assert close(out, ref, 1e-5) and np.all(out[0] == 0) and np.all(out[1] == 0) and np.all(out[4] == 0)
    g = np.ones_like(ref)
    gref = oracle.roi_align_backward(g, rois, data.shape, (2, 2), 1.0, 2)
    assert close(N(roi_align_backward(T(g), T(rois), data.shape, (2, 2), 1.0, 2)), gref, 1e-4)
    empty = roi_align_forward(T(data), T(rois[:0]), (2, 2), 1.0, 2)
    assert tuple(empty.shape) == (0, 3, 2, 2)
    z = N(roi_align_backward(T(g[:0]), T(rois[:0]), data.shape, (2, 2), 1.0, 2))
    assert z.shape == data.shape and np.all(z == 0)


def test_roi_align_autograd_and_linearity_cfg1_full_size():
    """BASELINE config 1 at full size: vs the C oracle, plus size-independent properties."""
    from mxdetection_b200.ops import ROIAlign, roi_align_forward
    c = syn.cfg1()
    data = T(c["data"]).requires_grad_(True)
    rois = T(c["rois"])
    out = ROIAlign(data, rois, c["pooled"], c["scale"], c["sample_ratio"])
    ref = cref.roi_align_forward(c["data"], c["rois"], c["pooled"], c["scale"], c["sample_ratio"])
    assert close(N(out), ref, 1e-5)
    out.backward(T(c["grad_out"]))
    gref = cref.roi_align_backward(c["grad_out"], c["rois"], c["data"].shape, c["pooled"], c["scale"], c["sample_ratio"])
    assert close(N(data.grad), gref, 1e-4)
    # linearity in the features and adjointness <RA(x), g> == <x, RA^T(g)>
    x2 = torch.randn_like(data)
    lhs = roi_align_forward(2.0 * data.detach() - 3.0 * x2, rois, c["pooled"], c["scale"], 2)
    rhs = 2.0 * out.detach() - 3.0 * roi_align_forward(x2, rois, c["pooled"], c["scale"], 2)
    assert torch.allclose(lhs, rhs, rtol=0, atol=2e-5)
    dot1 = (out.detach().double() * T(c["grad_out"]).double()).sum().item()
    dot2 = (data.detach().double() * data.grad.double()).sum().item()
    assert abs(dot1 - dot2) <= 1e-5 * max(1.0, abs(dot1))
    # constant map -> constant output wherever no sample is skipped
    const = roi_align_forward(torch.full_like(data.detach(), 1.5), rois, c["pooled"], c["scale"], 2)
    assert torch.all((const - 1.5).abs() <= 1e-6)


# ============================================================ levels + FPN (Spec G) ==
def test_map_roi_levels_bit_exact_incl_boundaries():
    from mxdetection_b200.models.roi_extractors import map_roi_levels
    rng = np.random.default_rng(9)
    r = np.concatenate([np.zeros((4000, 1)), rng.uniform(0, 600, (4000, 2)), rng.uniform(600, 1300, (4000, 2))], 1).astype(F)
    edge = []
    for s in (112.0, 224.0, 448.0):
        for d in (-1e-3, -1e-4, 0.0, 1e-4, 1e-3):
            v = np.float32(s + d)
            for _ in range(3):
                edge.append([0, 0, 0, v - 1, v - 1]); v = np.nextafter(v, np.float32(1e9), dtype=F)
    r = np.concatenate([r, np.asarray(edge, F)], 0)
    for L_ in (4, 5, 2, 1):
        assert np.array_equal(N(map_roi_levels(T(r), L_)), oracle.map_roi_levels(r, L_))
    assert np.array_equal(N(map_roi_levels(T(r[:, 1:].copy()), 4)), oracle.map_roi_levels(r[:, 1:], 4))


def test_fpn_roi_extractor_small_and_autograd():
    from mxdetection_b200.models.roi_extractors import SingleLevelRoI
    d = syn.fpn_roi_inputs(3, 2, 256, 320, 64, channels=16)
    ext = SingleLevelRoI(7, (4, 8, 16, 32), sample_num=2)
    feats = [T(f).requires_grad_(True) for f in d["feats"]]
    out = ext(feats, T(d["rois"]))
    lv = oracle.map_roi_levels(d["rois"], 4)
    assert len(set(lv.tolist())) >= 3
    ref = cref.roi_align_forward(d["feats"], d["rois"], (7, 7), d["scales"], 2, lv)
    assert close(N(out), ref, 1e-5)
    out.backward(T(d["grad_out"]))
    gref = cref.roi_align_backward(d["grad_out"], d["rois"], [f.shape for f in d["feats"]], (7, 7), d["scales"], 2, lv)
    for f, g in zip(feats, gref):
        assert close(N(f.grad), g, 1e-4)


def test_fpn_roi_stage_cfg3_shard_full_size(roi_path):
    """BASELINE config 3 geometry (800x1344, 4 levels, 256 ch, 512 RoIs/img), 2 images of the 8."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=2)
    lv = oracle.map_roi_levels(d["rois"], 4)
    out = roi_align_fpn_forward([T(f) for f in d["feats"]], T(d["rois"]), (7, 7), d["scales"], 2)   # levels in-kernel
    ref = cref.roi_align_forward(d["feats"], d["rois"], (7, 7), d["scales"], 2, lv)
    assert close(N(out), ref, 1e-5)
    g = roi_align_fpn_backward(T(d["grad_out"]), T(d["rois"]), [f.shape for f in d["feats"]], (7, 7), d["scales"], 2)
    gref = cref.roi_align_backward(d["grad_out"], d["rois"], [f.shape for f in d["feats"]], (7, 7), d["scales"], 2, lv)
    for a, b in zip(g, gref):
        assert close(N(a), b, 1e-4)


def test_fpn_roi_stage_cfg3_benchmarked_shape_all_8_images():
    """The shape bench.py times (BENCH / SCALE): all 8 images x 512 RoIs of BASELINE config 3 in ONE call - the unit
    counts, band / tile lists and work-stealing order of the full shard - forward and backward against the C port."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=8)
    assert d["rois"].shape[0] == 4096 and sum(f.nbytes for f in d["feats"]) == 731136000
    lv = oracle.map_roi_levels(d["rois"], 4)
    feats = [T(f) for f in d["feats"]]
    out = roi_align_fpn_forward(feats, T(d["rois"]), (7, 7), d["scales"], 2)
    ref = cref.roi_align_forward(d["feats"], d["rois"], (7, 7), d["scales"], 2, lv)
    assert close(N(out), ref, 1e-5)
    del feats, out, ref
    shapes = [f.shape for f in d["feats"]]
    g = roi_align_fpn_backward(T(d["grad_out"]), T(d["rois"]), shapes, (7, 7), d["scales"], 2)
    gref = cref.roi_align_backward(d["grad_out"], d["rois"], shapes, (7, 7), d["scales"], 2, lv)
    for a, b in zip(g, gref):
        assert close(N(a), b, 1e-4)
    # req=add on the same buffers: exactly twice the gradient (up to the accumulation-order bar)
    roi_align_fpn_backward(T(d["grad_out"]), T(d["rois"]), shapes, (7, 7), d["scales"], 2, grad_feats=g, accumulate=True)
    for a, b in zip(g, gref):
        assert close(N(a), 2 * b, 1e-4)


def test_fpn_roi_stage_is_deterministic_under_repetition():
    """Neither RoIAlign kernel of the benchmarked shape uses atomics (the RoIs of config 3 are clipped to the image, so the
    RED fallback list is empty): 40 back-to-back forward / backward launches on buffers that are re-used while the next
    launch's planners already run (programmatic dependent launch) must be BIT-identical.  A race between the producer
    warp, the progress words and the consumer warps of the persistent kernels would show up here as a flipped bit."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=8, with_features=False)
    shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    rois = T(d["rois"])
    gout = torch.randn((rois.shape[0], 256, 7, 7), device="cuda", generator=g)
    out0 = roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2).clone()
    out = torch.empty_like(out0)
    for _ in range(40):
        out.fill_(float("nan"))
        roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
        assert torch.equal(out, out0)
    del feats
    g0 = [t.clone() for t in roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2)]
    grads = [torch.empty_like(t) for t in g0]
    for _ in range(40):
        for t in grads:
            t.fill_(float("nan"))
        roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)
        assert all(torch.equal(a, b) for a, b in zip(grads, g0))


@pytest.mark.parametrize("W,C", [(400, 20), (512, 8), (520, 8), (352, 6), (356, 4)])
def test_roi_align_ring_wide_maps_and_class_edges(W, C, roi_path):
    """Row pitch classes of the ring forward: 352 / 704 / 1408 / 2048 bytes.  W = 352 is the widest 2-channel class,
    356..512 run one channel per unit, 520 exceeds every class (the band kernels take the call); tall maps wrap the ring."""
    from mxdetection_b200.ops import roi_align_forward
    rng = np.random.default_rng(W + C)
    H = 150
    data = rng.standard_normal((2, C, H, W)).astype(F)
    R = 120
    x1 = rng.uniform(-10, W * 4 - 20, R); y1 = rng.uniform(-10, H * 4 - 20, R)
    bw = np.exp(rng.uniform(np.log(8), np.log(300), R)); bh = np.exp(rng.uniform(np.log(8), np.log(500), R))
    rois = np.stack([rng.integers(0, 2, R), x1, y1, x1 + bw, y1 + bh], 1).astype(F)
    rois[:4] = [[0, 0, 0, W * 4 - 1, 40], [1, W * 4 - 60, H * 4 - 60, W * 4 + 5, H * 4 + 5], [0, 3, 500, 200, 599.5], [1, 0, 0, 30, H * 4 - 1]]
    out = N(roi_align_forward(T(data), T(rois), (7, 7), 0.25, 2))
    assert close(out, cref.roi_align_forward(data, rois, (7, 7), 0.25, 2), 1e-5)


def test_roi_align_crowded_band_and_tile(roi_path):
    """1200 RoIs in one corner of one image (600 of them identical): one band / one tile carries far more RoIs than a
    shared-memory table chunk or a message batch holds; the identical RoIs make the backward accumulate 600 times
    into the same pixels (order-dependent rounding: 1e-4 relative still holds)."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    rng = np.random.default_rng(51)
    shapes = [(2, 32, 64, 80), (2, 32, 32, 40), (2, 32, 16, 20), (2, 32, 8, 10)]
    scales = [0.25, 0.125, 0.0625, 0.03125]
    feats = [rng.normal(0, 1, s).astype(F) for s in shapes]
    xy = rng.uniform(0, 60, (600, 2)); wh = rng.uniform(8, 40, (600, 2))
    crowd = np.concatenate([np.zeros((600, 1)), xy, xy + wh], 1)
    same = np.tile(np.array([[0, 12.5, 9.25, 47.0, 39.5]]), (600, 1))
    rois = np.concatenate([crowd, same, np.array([[1, 100, 100, 180, 170]])]).astype(F)
    lv = oracle.map_roi_levels(rois, 4)
    out = roi_align_fpn_forward([T(f) for f in feats], T(rois), (7, 7), scales, 2)
    assert close(N(out), cref.roi_align_forward(feats, rois, (7, 7), scales, 2, lv), 1e-5)
    gout = rng.normal(0, 1, (rois.shape[0], 32, 7, 7)).astype(F)
    g = roi_align_fpn_backward(T(gout), T(rois), shapes, (7, 7), scales, 2)
    gref = cref.roi_align_backward(gout, rois, shapes, (7, 7), scales, 2, lv)
    for x, r in zip(g, gref):
        assert np.all(np.abs(N(x) - r) <= 1e-4 * np.maximum(1.0, np.abs(r)) + 1e-4 * np.abs(r).max() * 0.01)


def test_host_roi_stage_pipeline_matches_device_path():
    """Host-buffer entry (H2D | fwd+bwd | D2H pipelined per image over three streams) == the plain device ops."""
    from mxdetection_b200.ops import HostRoIStage, roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.fpn_roi_inputs(3, 5, 256, 320, 64, channels=32)          # 5 images, 64 RoIs each
    feats_h = [torch.from_numpy(f).pin_memory() for f in d["feats"]]
    rois_h = torch.from_numpy(d["rois"]).pin_memory(); gout_h = torch.from_numpy(d["grad_out"]).pin_memory()
    out_h = torch.empty(d["grad_out"].shape).pin_memory()
    grads_h = [torch.empty(f.shape).pin_memory() for f in d["feats"]]
    ref = roi_align_fpn_forward([T(f) for f in d["feats"]], T(d["rois"]), (7, 7), d["scales"], 2)
    gref = roi_align_fpn_backward(T(d["grad_out"]), T(d["rois"]), [f.shape for f in d["feats"]], (7, 7), d["scales"], 2)
    for split in (1, 2):                                             # whole images / two channel slices per image
        stage = HostRoIStage([f.shape for f in d["feats"]], 64, (7, 7), d["scales"], 2, DEV, depth=2, channel_split=split)
        for _ in range(2):                                           # second call re-uses the slots
            out_h.zero_(); [g.zero_() for g in grads_h]
            stage.forward_backward(feats_h, rois_h, gout_h, out_h, grads_h).synchronize()
        assert close(out_h.numpy(), N(ref), 1e-6)
        for a, b in zip(grads_h, gref):
            assert close(a.numpy(), N(b), 1e-4)


def test_mask_branch_cfg4_14x14():
    from mxdetection_b200.ops import roi_align_fpn_forward
    d = syn.cfg4_mask(batch=1)
    lv = oracle.map_roi_levels(d["rois"], 4)
    out = roi_align_fpn_forward([T(f) for f in d["feats"]], T(d["rois"]), (14, 14), d["scales"], 2, levels=T(lv))
    ref = cref.roi_align_forward(d["feats"], d["rois"], (14, 14), d["scales"], 2, lv)
    assert close(N(out), ref, 1e-5)
    from mxdetection_b200.ops import roi_align_fpn_backward
    g = roi_align_fpn_backward(T(d["grad_out"]), T(d["rois"]), [f.shape for f in d["feats"]], (14, 14), d["scales"], 2)
    gref = cref.roi_align_backward(d["grad_out"], d["rois"], [f.shape for f in d["feats"]], (14, 14), d["scales"], 2, lv)
    for a, b in zip(g, gref):
        assert close(N(a), b, 1e-4)


def test_host_roi_stage_images_without_rois():
    """Images 1 and 3 of 5 carry no RoI: their gradient planes come back as zeros, the rest matches the device ops."""
    from mxdetection_b200.ops import HostRoIStage, roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.fpn_roi_inputs(3, 5, 256, 320, 64, channels=32)
    keep = ~np.isin(d["rois"][:, 0].astype(int), (1, 3))
    rois, gout = d["rois"][keep], d["grad_out"][keep]
    feats_h = [torch.from_numpy(f).pin_memory() for f in d["feats"]]
    rois_h = torch.from_numpy(rois).pin_memory(); gout_h = torch.from_numpy(gout).pin_memory()
    out_h = torch.full(gout.shape, 5.0).pin_memory()
    grads_h = [torch.full(f.shape, 5.0).pin_memory() for f in d["feats"]]
    stage = HostRoIStage([f.shape for f in d["feats"]], 64, (7, 7), d["scales"], 2, DEV, depth=2)
    stage.forward_backward(feats_h, rois_h, gout_h, out_h, grads_h).synchronize()
    ref = roi_align_fpn_forward([T(f) for f in d["feats"]], T(rois), (7, 7), d["scales"], 2)
    gref = roi_align_fpn_backward(T(gout), T(rois), [f.shape for f in d["feats"]], (7, 7), d["scales"], 2)
    assert close(out_h.numpy(), N(ref), 1e-6)
    for a, b in zip(grads_h, gref):
        assert close(a.numpy(), N(b), 1e-4) and float(a[1].abs().max()) == 0.0 and float(a[3].abs().max()) == 0.0


# ================================================================ top-k (Spec B/H) ==
# (40,40) .. (130,130): the sort's register phase alone / plus one and two shared-memory merge sizes; (8192,8192): full capacity
@pytest.mark.parametrize("n,k", [(1, 1), (5, 10), (63, 7), (40, 40), (64, 64), (100, 100), (130, 130), (2048, 2048),
                                  (8192, 2000), (8192, 8192), (8193, 2000), (50000, 2000),
                                  (217413, 2000), (201600, 6000), (300000, 1)])
def test_topk_stable_bit_exact(n, k):
    from mxdetection_b200.ops import topk_stable
    rng = np.random.default_rng(n + k)
    z = rng.normal(-4, 2, (3, n))
    s = (1 / (1 + np.exp(-z))).astype(F)
    s[1] = np.round(s[1] * 64) / 64                # heavy ties
    s[2, : n // 2] = 0.0; s[2, n // 2] = -0.0      # zeros of both signs compare equal -> index decides
    idx, vals = topk_stable(T(s), k)
    for r in range(3):
        ref = oracle.topk_stable(s[r], k)
        assert np.array_equal(N(idx)[r], ref), "segment %d" % r
        assert np.array_equal(N(vals)[r], s[r][ref])


def test_topk_adversarial_inputs_take_the_exact_fallback():
    from mxdetection_b200.ops import topk_stable
    rng = np.random.default_rng(77)
    n = 4096 * 8
    s = rng.uniform(0, 0.5, n).astype(F)
    hot = (np.arange(n) % 4096) < 2000             # > 8192 rows beat the group-maxima bound
    s[hot] = rng.uniform(0.5, 1.0, hot.sum()).astype(F)
    const = np.full(n, 0.25, F)                    # all equal: the index alone orders
    desc = np.linspace(1, 0, n).astype(F)
    for arr in (s, const, desc, desc[::-1].copy()):
        idx, _ = topk_stable(T(arr), 2000)
        assert np.array_equal(N(idx), oracle.topk_stable(arr, 2000))


# ===================================================================== NMS (Spec B) ==
@pytest.mark.parametrize("name", ["nms_a", "nms_b", "nms_c"])
def test_nms_golden(golden_dir, name):
    from mxdetection_b200.ops import nms_indices
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    keep, num = nms_indices(T(g["boxes"]), T(g["scores"]), float(g["thr"]), delta=0.0)
    k = int(num.item())
    assert np.array_equal(N(keep)[:k], g["keep"]) and np.all(N(keep)[k:] == -1)


# 7000 boxes: four-buffer band ring of the scan kernel; 8192 (the in-CTA sort capacity): three buffers
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 128, 129, 1000, 2000, 4100, 7000, 8192])
@pytest.mark.parametrize("delta", [0.0, 1.0])
def test_nms_vs_oracle_bit_exact(n, delta):
    from mxdetection_b200.ops import nms_indices
    rng = np.random.default_rng(n * 3 + int(delta))
    span = 40 * math.sqrt(n) + 20
    xy = rng.uniform(0, span, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    boxes[n // 3] = boxes[0]                                     # exact duplicates
    scores = (np.round(rng.uniform(0, 1, n) * 40) / 40).astype(F)   # ties
    ref = oracle.nms(boxes, scores, 0.7, delta=delta)
    keep, num = nms_indices(T(boxes), T(scores), 0.7, delta=delta)
    assert int(num.item()) == len(ref) and np.array_equal(N(keep)[: len(ref)], ref)


@pytest.mark.parametrize("n,k", [(9000, 9000), (20000, 12000), (70001, 70001), (300, 300), (16384, 8193)])
def test_topk_stable_beyond_the_in_cta_sort(n, k):
    """k above MXD_SORT_CAP (8192): the chunk-sort + rank-merge path (also forced on a short input)."""
    from mxdetection_b200.ops import topk_stable
    rng = np.random.default_rng(n)
    s = (np.round(rng.uniform(0, 1, (2, n)) * 4096) / 4096).astype(F)     # ties
    s[1, ::3] = 0.125
    if k <= 8192:
        os.environ["MXD_TOPK_FORCE_LONG"] = "1"      # read by the workspace query and by the launcher
    try:
        idx, vals = topk_stable(T(s), k)
    finally:
        os.environ.pop("MXD_TOPK_FORCE_LONG", None)
    for r in range(2):
        ref = oracle.topk_stable(s[r], k)
        assert np.array_equal(N(idx)[r], ref) and np.array_equal(N(vals)[r], s[r][ref])


@pytest.mark.parametrize("n,force", [(9000, False), (12000, False), (20000, False), (700, True), (64, True)])
def test_nms_long_segments_beyond_8192_rows(n, force):
    """mx.nd.contrib.box_nms has no row cap: above 8192 rows the sort takes the chunked path, and above ~9400 rows the
    greedy resolve reads the row-major mask from L2 (forced here for short inputs too)."""
    from mxdetection_b200.ops import nms_indices, box_nms
    rng = np.random.default_rng(n)
    span = 30 * math.sqrt(n) + 20
    xy = rng.uniform(0, span, (n, 2)); wh = rng.uniform(4, 100, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = (np.round(rng.uniform(0, 1, n) * 512) / 512).astype(F)
    if force:
        os.environ["MXD_NMS_FORCE_GLOBAL"] = "1"
    try:
        ref = oracle.nms(boxes, scores, 0.6, delta=0.0)
        keep, num = nms_indices(T(boxes), T(scores), 0.6, delta=0.0)
        assert int(num.item()) == len(ref) and np.array_equal(N(keep)[: len(ref)], ref)
        ref2 = oracle.nms(boxes, scores, 0.6, delta=0.0, max_out=37, valid_thresh=0.3)
        keep, num = nms_indices(T(boxes), T(scores), 0.6, delta=0.0, max_out=37, valid_thresh=0.3)
        assert int(num.item()) == len(ref2) and np.array_equal(N(keep)[: len(ref2)], ref2)
        if n <= 12000:
            data = np.concatenate([rng.integers(0, 3, (n, 1)).astype(F), scores[:, None], boxes], 1)[None]
            out = N(box_nms(T(data), overlap_thresh=0.6, valid_thresh=0.0, id_index=0))
            assert np.array_equal(out, oracle.box_nms_mx(data, overlap_thresh=0.6, valid_thresh=0.0, id_index=0))
    finally:
        os.environ.pop("MXD_NMS_FORCE_GLOBAL", None)


def test_nms_batched_ragged_segments():
    """mxd_nms_batched: segments of different lengths (empty, 1, > 64, > 8192 rows) of one box array; keep lists are
    global row indices and equal the per-segment oracle."""
    from mxdetection_b200.ops import nms_batched
    rng = np.random.default_rng(12)
    lens = [0, 1, 700, 65, 0, 2000, 9001, 3]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(off[-1])
    xy = rng.uniform(0, 900, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = (np.round(rng.uniform(0, 1, n) * 256) / 256).astype(F)
    ids = rng.integers(0, 4, n).astype(np.int32)
    for kw in (dict(), dict(topk=500, max_out=100), dict(ids=ids, force_suppress=False, valid_thresh=0.25)):
        tk = {k: (T(v) if k == "ids" else v) for k, v in kw.items()}
        keep, num = nms_batched(T(boxes), T(scores), T(off), max(lens), 0.5, delta=1.0, **tk)
        keep, num = N(keep), N(num)
        for s_, (a, b) in enumerate(zip(off[:-1], off[1:])):
            okw = dict(kw)
            if "ids" in okw:
                okw["ids"] = ids[a:b]
            ref = oracle.nms(boxes[a:b], scores[a:b], 0.5, delta=1.0, **okw) + a
            assert num[s_] == len(ref) and np.array_equal(keep[s_, : len(ref)], ref), (kw, s_)
            assert np.all(keep[s_, len(ref):] == -1)


def test_nms_dependency_chains():
    """Worst case of the scan kernel's fixed point: every box overlaps only its neighbours, so whether box i survives
    depends on box i-1, ... all the way down (64 rounds per block), across block boundaries, with ties."""
    from mxdetection_b200.ops import nms_indices
    for n, step in ((64, 10.0), (200, 10.0), (1000, 7.0), (333, 12.0)):
        x = np.arange(n, dtype=np.float64) * step
        boxes = np.stack([x, np.zeros(n), x + 15.0, np.full(n, 10.0)], 1).astype(F)
        for scores in (np.linspace(1, 0.1, n).astype(F),            # left to right
                       np.linspace(0.1, 1, n).astype(F),             # right to left
                       np.full(n, 0.5, F)):                          # all ties: index order
            ref = oracle.nms(boxes, scores, 0.1, delta=0.0)
            keep, num = nms_indices(T(boxes), T(scores), 0.1, delta=0.0)
            assert int(num.item()) == len(ref) and np.array_equal(N(keep)[: len(ref)], ref), (n, step)


def test_nms_options_topk_validthresh_ids_maxout():
    from mxdetection_b200.ops import nms_indices, nms
    rng = np.random.default_rng(4)
    n = 900
    xy = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(4, 90, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = rng.uniform(0, 1, n).astype(F)
    ids = rng.integers(0, 5, n).astype(np.int32)
    for kw in (dict(topk=200), dict(valid_thresh=0.4), dict(max_out=50), dict(ids=ids, force_suppress=False),
               dict(ids=ids, force_suppress=True), dict(topk=300, valid_thresh=0.2, max_out=70, ids=ids, force_suppress=False)):
        ref = oracle.nms(boxes, scores, 0.5, delta=0.0, **kw)
        tk = {k: (T(v) if k == "ids" else v) for k, v in kw.items()}
        keep, num = nms_indices(T(boxes), T(scores), 0.5, delta=0.0, **tk)
        assert int(num.item()) == len(ref) and np.array_equal(N(keep)[: len(ref)], ref), kw
    dets = np.concatenate([boxes, scores[:, None]], 1)
    kept, inds = nms(T(dets), 0.5)
    ref = oracle.nms(boxes, scores, 0.5, delta=1.0)
    assert np.array_equal(N(inds), ref) and np.array_equal(N(kept), dets[ref])
    # idempotence: NMS of the survivors keeps all of them
    k2, n2 = nms_indices(T(boxes[ref]), T(scores[ref]), 0.5, delta=1.0)
    assert int(n2.item()) == len(ref)
    empty, ne = nms_indices(T(boxes[:0]), T(scores[:0]), 0.5)
    assert int(ne.item()) == 0


def test_box_nms_mx_tensor_api_and_backward():
    from mxdetection_b200.ops import box_nms, box_nms_backward
    x = np.array([[0, .5, .1, .1, .2, .2], [1, .4, .1, .1, .2, .2], [0, .3, .1, .1, .14, .14], [2, .6, .5, .5, .7, .8]], F)
    out = N(box_nms(T(x), overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=True))
    exp = np.array([[2, .6, .5, .5, .7, .8], [0, .5, .1, .1, .2, .2], [-1] * 6, [-1] * 6], F)
    assert np.array_equal(out, exp)                                             # KAT-1
    rng = np.random.default_rng(12)
    B, Nn = 3, 500
    xy = rng.uniform(0, 1, (B, Nn, 2)); wh = rng.uniform(0.02, 0.3, (B, Nn, 2))
    data = np.concatenate([rng.integers(0, 4, (B, Nn, 1)), rng.uniform(-0.2, 1, (B, Nn, 1)), xy, xy + wh], 2).astype(F)
    for kw in (dict(), dict(topk=100), dict(id_index=0), dict(id_index=0, force_suppress=True), dict(valid_thresh=0.3),
               dict(out_format="center"), dict(in_format="center", out_format="corner")):
        ref, ridx = oracle.box_nms_mx(data, overlap_thresh=0.45, coord_start=2, score_index=1, return_index=True, **kw)
        got, gidx = box_nms(T(data), overlap_thresh=0.45, coord_start=2, score_index=1, return_index=True, **kw)
        assert np.array_equal(N(gidx), ridx), kw
        assert np.array_equal(N(got), ref), kw
    og = rng.standard_normal(data.shape).astype(F)
    ig = N(box_nms_backward(T(og), gidx))
    exp = np.zeros_like(og)
    for b in range(B):
        for r in range(Nn):
            if ridx[b, r] >= 0:
                exp[b, ridx[b, r]] = og[b, r]
    assert np.array_equal(ig, exp)


# ========================================================= anchors / IoU / assigner ==
def test_anchors_and_flags_bit_exact():
    from mxdetection_b200.core.anchor import AnchorGenerator, anchor_inside_flags
    for base, stride, (fh, fw) in [(4, 4, (50, 68)), (16, 16, (13, 17)), (64, 64, (3, 5))]:
        ag = AnchorGenerator(base, [8], [0.5, 1.0, 2.0])
        ref = oracle.grid_anchors(oracle.gen_base_anchors(base, [8], [0.5, 1, 2]), fh, fw, stride)
        got = ag.grid_anchors((fh, fw), stride)
        assert np.array_equal(N(got), ref)
        v = ag.valid_flags((fh, fw), (fh - 1, fw - 2))
        assert np.array_equal(N(v), oracle.valid_flags(fh, fw, fh - 1, fw - 2, 3))
        for ab in (0, 8, -1):
            ins = anchor_inside_flags(got, v, (fh * stride - 3, fw * stride - 5), ab)
            assert np.array_equal(N(ins), oracle.inside_flags(ref, N(v), fh * stride - 3, fw * stride - 5, ab))


def test_bbox_overlaps_bit_exact():
    from mxdetection_b200.core.bbox import bbox_overlaps
    rng = np.random.default_rng(2)
    g = syn.gt_boxes(rng, 800, 1344, 57); a = syn.gt_boxes(rng, 800, 1344, 3001)
    a[5] = g[3]
    for d in (1.0, 0.0):
        assert np.array_equal(N(bbox_overlaps(T(g), T(a), delta=d)), oracle.bbox_overlaps(g, a, d))


def test_assigner_kat4_and_random_bit_exact():
    from mxdetection_b200.core.bbox import MaxIoUAssigner
    gts = np.array([[0, 0, 9, 9], [10, 0, 19, 9], [100, 100, 149, 149]], F)
    anchors = np.array([[100, 100, 149, 149], [300, 300, 310, 310], [5, 0, 14, 9], [100, 100, 149, 124], [0, 0, 9, 4],
                        [400, 0, 409, 9]], F)
    res = MaxIoUAssigner(0.7, 0.3, 0.3).assign(T(anchors), T(gts), T(np.array([7, 8, 9], np.int32)))
    assert N(res.gt_inds).tolist() == [3, 0, 2, -1, 1, 0] and N(res.labels).tolist() == [9, 0, 8, 0, 7, 0]
    keep = [0, 1, 2, 3, 5]
    res = MaxIoUAssigner(0.7, 0.3, 0.3).assign(T(anchors[keep]), T(gts))
    assert N(res.gt_inds).tolist() == [3, 0, 2, -1, 0]
    # random, batched, ragged GT counts (incl. 0), flags, both threshold sets
    rng = np.random.default_rng(31)
    base = oracle.gen_base_anchors(16, [4, 8], [0.5, 1, 2])
    anc = oracle.grid_anchors(base, 25, 38, 16)
    flags = oracle.inside_flags(anc, np.ones(len(anc), np.uint8), 400, 600, 0)
    gl = [syn.gt_boxes(rng, 400, 600, g) for g in (30, 1, 0, 300)]      # 300 > one smem chunk of 256
    G, num = syn.padded_gts(gl, 300)
    labels = rng.integers(1, 81, (4, 300)).astype(np.int32)
    for thr in ((0.7, 0.3, 0.3), (0.5, 0.5, 0.5), (0.5, 0.5, 0.0)):
        a, m, l = MaxIoUAssigner(*thr).assign_batch(T(anc), T(G), T(num), T(labels), T(flags))
        for b in range(4):
            ra, rm, rl = oracle.max_iou_assign(anc, gl[b], labels[b, : num[b]], *thr, flags=flags)
            assert np.array_equal(N(a)[b], ra), (thr, b)
            assert np.array_equal(N(m)[b], rm) and np.array_equal(N(l)[b], rl)


def test_assigner_degenerate_boxes_and_touching_edges():
    """Quick-reject window vs the exact Spec D test: zero-area boxes (delta 0), boxes that only touch
    (iw == 0 / iw == delta), far-apart boxes, identical boxes, huge coordinates - labels stay bit-exact."""
    from mxdetection_b200.core.bbox import MaxIoUAssigner
    rng = np.random.default_rng(77)
    for delta in (1.0, 0.0):
        gts = syn.gt_boxes(rng, 600, 800, 40)
        gts[3] = [50, 60, 50, 80]                  # zero width (area 0 when delta = 0)
        gts[4] = [200, 200, 260, 200]              # zero height
        gts[5] = [1.6e7, 1.6e7, 1.6e7 + 64, 1.6e7 + 64]     # spacing of fp32 is 1-2 px here
        anchors = syn.gt_boxes(rng, 600, 800, 5000)
        anchors[:40] = gts                          # exact copies -> IoU 1 ties
        anchors[40:80, 0] = gts[:, 2]; anchors[40:80, 2] = gts[:, 2] + 30     # touch the right edge: iw = delta
        anchors[80:120, 2] = gts[:, 0] - 1; anchors[80:120, 0] = gts[:, 0] - 25   # one pixel to the left: iw = delta - 1
        anchors[120] = [10, 10, 10, 10]            # zero-area anchor, not on any zero-area GT
        anchors[121] = [1.6e7 + 1, 1.6e7 + 1, 1.6e7 + 40, 1.6e7 + 40]
        a = MaxIoUAssigner(0.5, 0.4, 0.2, delta=delta).assign(T(anchors), T(gts))
        ra, rm, _ = oracle.max_iou_assign(anchors, gts, None, 0.5, 0.4, 0.2, delta=delta)
        # 0/0 pairs (Spec D): NaN propagates through max_overlaps and gt_max exactly as in the oracle (numpy max / argmax)
        if delta == 0.0:
            assert np.isnan(rm).any()
        assert np.array_equal(N(a.gt_inds), ra) and np.array_equal(N(a.max_overlaps), rm, equal_nan=True)
        rc = cref.max_iou_assign_batch(anchors, gts[None], np.array([len(gts)], np.int32), None, None, 0.5, 0.4, 0.2, delta)
        assert np.array_equal(rc[0][0], ra) and np.array_equal(rc[1][0], rm, equal_nan=True)     # C port == NumPy oracle


def test_assigner_cfg4_full_size_vs_c_oracle():
    """BASELINE config 4b: 268 569 FPN anchors x 100 GTs, batch 2."""
    from mxdetection_b200.core.anchor import AnchorGenerator, anchor_inside_flags, anchor_assign
    d = syn.assigner_inputs(4, 2)
    anchors, valid = [], []
    for (fh, fw), s in zip(d["feat_shapes"], d["strides"]):
        ag = AnchorGenerator(s, [8], [0.5, 1.0, 2.0])
        anchors.append(ag.grid_anchors((fh, fw), s)); valid.append(ag.valid_flags((fh, fw), (fh, fw)))
    anchors = torch.cat(anchors); valid = torch.cat(valid)
    assert anchors.shape[0] == 268569
    inside = anchor_inside_flags(anchors, valid, d["img_shape"], 0)
    a, m, l = anchor_assign(anchors, inside, T(d["gts"]), T(d["num_gts"]), T(d["gt_labels"]))
    ra, rm, rl = cref.max_iou_assign_batch(N(anchors), d["gts"], d["num_gts"], d["gt_labels"], N(inside))
    assert np.array_equal(N(a), ra) and np.array_equal(N(m), rm) and np.array_equal(N(l), rl)
    assert (ra > 0).sum() >= 200 and (ra == -1).sum() > 0


# ================================================================== codec (Spec F) ==
def test_codec_vs_oracle():
    from mxdetection_b200.core.bbox import bbox2delta, delta2bbox
    rng = np.random.default_rng(8)
    p = syn.gt_boxes(rng, 800, 1344, 5000)
    dl = rng.normal(0, 0.5, (5000, 4)).astype(F); dl[:50, 2:] *= 20
    for stds, shape in [((1, 1, 1, 1), (800, 1344)), ((.1, .1, .2, .2), None)]:
        ref = oracle.delta2bbox(p, dl, stds=stds, max_shape=shape)
        got = N(delta2bbox(T(p), T(dl), stds=stds, max_shape=shape))
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
        assert (got == ref).mean() > 0.9999
    g = p + rng.uniform(-6, 6, p.shape).astype(F)
    ref = oracle.bbox2delta(p, g, stds=(.1, .1, .2, .2))
    got = N(bbox2delta(T(p), T(g), stds=(.1, .1, .2, .2)))
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6) and (got == ref).mean() > 0.9999


# ========================================================== RPN proposals (Spec H) ==
def _run_rpn(d, cfgkw):
    from mxdetection_b200.models.rpn_heads import RPNHead, ProposalConfig
    head = RPNHead()
    cfg = ProposalConfig(**cfgkw)
    props, nv, handle = head.get_proposals([T(s) for s in d["scores"]], [T(x) for x in d["deltas"]], d["feat_shapes"],
                                           d["img_shapes"], cfg, return_workspace=True)
    B = d["scores"][0].shape[0]
    stages = RPNHead.stages(handle, B)
    return N(props), N(nv), [N(s) for s in stages]


def _check_rpn_stagewise(d, cfgkw, props, nv, stages):
    idx, boxes, keep, counts = stages
    B = d["scores"][0].shape[0]
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    nms_pre, nms_post, max_num = cfgkw["nms_pre"], cfgkw["nms_post"], cfgkw["max_num"]
    msz = cfgkw.get("min_bbox_size", 0)
    for b in range(B):
        cat = []
        for l, (fh, fw) in enumerate(d["feat_shapes"]):
            s = d["scores"][l][b]
            ref_idx = oracle.topk_stable(s, nms_pre)
            k = len(ref_idx)
            assert counts[b, l, 0] == k
            assert np.array_equal(idx[b, l, :k], ref_idx), "top-k indices must be bit-exact"       # stage 1
            anc = oracle.grid_anchors(base[l], fh, fw, d["strides"][l])[ref_idx]
            ref_boxes = oracle.delta2bbox(anc, d["deltas"][l][b][ref_idx], max_shape=tuple(d["img_shapes"][b]))
            gb = boxes[b, l, :k]
            assert np.abs(gb - ref_boxes).max() <= 1e-5 * 2000                                      # stage 2 (tolerance)
            valid = None
            if msz > 0:
                valid = ((gb[:, 2] - gb[:, 0] + 1) >= msz) & ((gb[:, 3] - gb[:, 1] + 1) >= msz)
            # stage 3: oracle NMS on the GPU-decoded boxes must give the GPU keep list, bit-exact
            ref_keep = oracle.nms(gb, s[ref_idx], cfgkw["nms_thr"], delta=1.0, valid_mask=valid, max_out=nms_post)
            nk = counts[b, l, 1]
            assert nk == len(ref_keep) and np.array_equal(keep[b, l, :nk], ref_keep)
            cat.append(np.concatenate([gb[ref_keep], s[ref_idx][ref_keep, None]], 1))
        cat = np.concatenate(cat, 0)
        if len(cat) > max_num:                                                                      # stage 4
            cat = cat[stable_order_desc(cat[:, 4])[:max_num]]
        assert nv[b] == len(cat)
        assert np.array_equal(props[b, : len(cat)], cat) and np.all(props[b, len(cat):] == 0)


@pytest.mark.parametrize("cfgkw", [dict(nms_pre=300, nms_post=100, max_num=150, nms_thr=0.7),
                                   dict(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7),
                                   dict(nms_pre=500, nms_post=500, max_num=4000, nms_thr=0.5, min_bbox_size=20)])
def test_rpn_proposals_small_stagewise_and_end_to_end(cfgkw):
    d = syn.rpn_inputs(2, 3, 160, 224)
    props, nv, stages = _run_rpn(d, cfgkw)
    _check_rpn_stagewise(d, cfgkw, props, nv, stages)
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    ro, rn = oracle.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfgkw)
    assert np.array_equal(nv, rn)
    assert np.abs(props - ro).max() <= 1e-5 * 2000 and (props == ro).mean() > 0.9999


def test_rpn_proposals_degenerate_scores_and_tiny_maps():
    """All scores equal (the index alone orders every level), heavy quantisation, a 32x32 image (1x1 coarse maps,
    fewer anchors than nms_pre on every level), huge deltas (the ratio clip and the image clip decide everything)."""
    rng = np.random.default_rng(5)
    for (ih, iw), mode in (((160, 224), "const"), ((160, 224), "quant"), ((32, 32), "rand"), ((96, 64), "bigdelta")):
        d = syn.rpn_inputs(7, 2, ih, iw)
        for l in range(len(d["scores"])):
            if mode == "const":
                d["scores"][l][:] = 0.5
            elif mode == "quant":
                d["scores"][l] = (np.round(d["scores"][l] * 8) / 8).astype(F)
            elif mode == "bigdelta":
                d["deltas"][l] = rng.normal(0, 3.0, d["deltas"][l].shape).astype(F)
        for cfgkw in (dict(nms_pre=300, nms_post=100, max_num=150, nms_thr=0.7), dict(nms_pre=50, nms_post=50, max_num=20, nms_thr=0.3, min_bbox_size=8)):
            props, nv, stages = _run_rpn(d, cfgkw)
            base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
            ro, rn = oracle.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfgkw)
            assert np.array_equal(nv, rn), (mode, cfgkw)
            assert np.abs(props - ro).max() <= 1e-5 * 2000 and (props == ro).mean() > 0.999, (mode, cfgkw)


def test_rpn_proposals_cfg2_full_size():
    """BASELINE config 2: 800x1088, 217 413 anchors/img, top-2000/level, NMS 0.7, post 1000, batch 2."""
    cfgkw = dict(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
    d = syn.cfg2(batch=2)
    assert sum(s.shape[1] for s in d["scores"]) == 217413
    props, nv, stages = _run_rpn(d, cfgkw)
    assert stages[3][:, :, 0].sum() == 2 * 8663
    _check_rpn_stagewise(d, cfgkw, props, nv, stages)
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    ro, rn = cref.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfgkw)
    assert np.array_equal(nv, rn) and np.array_equal(props, ro)
    assert np.all(props[:, :-1, 4] >= props[:, 1:, 4])           # sortedness of the truncated output


def test_two_host_threads_on_two_streams_match_serial_results():
    """include/mxdet.h: 're-entrant; concurrent calls on distinct streams / workspaces are safe'.  Two host threads, each on
    its own stream (hence its own workspaces), interleave RoIAlign forward / backward and the proposal stage; every result
    must equal the one computed serially (all three are bit-reproducible)."""
    import threading
    from mxdetection_b200.models.rpn_heads import ProposalConfig, RPNHead
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=2)
    feats = [T(f) for f in d["feats"]]; rois = T(d["rois"]); gout = T(d["grad_out"])
    shapes = [f.shape for f in d["feats"]]
    r = syn.rpn_inputs(31, 2, 800, 1088)
    sc = [T(x) for x in r["scores"]]; dl = [T(x) for x in r["deltas"]]; shp = T(r["img_shapes"])
    cfg = ProposalConfig(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
    ref_out = roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2).clone()
    ref_g = [t.clone() for t in roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2)]
    ref_p, ref_n = [t.clone() for t in RPNHead().get_proposals(sc, dl, r["feat_shapes"], shp, cfg)]
    torch.cuda.synchronize()
    errors = []

    def worker(order):
        try:
            st = torch.cuda.Stream()
            head = RPNHead()
            with torch.cuda.stream(st):
                for it in range(6):
                    for what in order:
                        if what == "f":
                            ok = torch.equal(roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2), ref_out)
                        elif what == "b":
                            g = roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2)
                            ok = all(torch.equal(a, b) for a, b in zip(g, ref_g))
                        else:
                            p, n = head.get_proposals(sc, dl, r["feat_shapes"], shp, cfg)
                            ok = torch.equal(p, ref_p) and torch.equal(n, ref_n)
                        if not ok:
                            errors.append("thread %s iteration %d: %s differs" % (order, it, what))
                st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(o,)) for o in ("fbp", "pbf")]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors


def test_rpn_proposals_and_assigner_repeatable_bit_for_bit():
    """The latency-bound stages use shared-memory atomics for compaction and per-GT maxima; their RESULTS must not depend
    on arrival order: 25 back-to-back launches of the config-5-shard proposal stage (8 images, 4-CTA clusters) and of
    the assigner are bit-identical (a race in the cluster top-k, the NMS ring or pass 1 / pass 2 would show up here)."""
    from mxdetection_b200.models.rpn_heads import ProposalConfig, RPNHead
    from mxdetection_b200.core.bbox import MaxIoUAssigner
    d = syn.rpn_inputs(77, 8, 800, 1344)
    sc = [T(x) for x in d["scores"]]; dl = [T(x) for x in d["deltas"]]; shp = T(d["img_shapes"])
    head, cfg = RPNHead(), ProposalConfig(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
    p0, n0 = head.get_proposals(sc, dl, d["feat_shapes"], shp, cfg)
    p0, n0 = p0.clone(), n0.clone()
    for _ in range(25):
        p, n = head.get_proposals(sc, dl, d["feat_shapes"], shp, cfg)
        assert torch.equal(n, n0) and torch.equal(p, p0)
    rng = np.random.default_rng(4)
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    anc = np.concatenate([oracle.grid_anchors(b, fh, fw, st) for b, (fh, fw), st in zip(base, d["feat_shapes"], d["strides"])])
    gts = syn.gt_boxes(rng, 800, 1344, 100)
    asg = MaxIoUAssigner(0.7, 0.3, 0.3)
    a0 = asg.assign(T(anc), T(gts))
    a0 = [t.clone() for t in (a0.gt_inds, a0.max_overlaps)]
    for _ in range(25):
        a = asg.assign(T(anc), T(gts))
        assert torch.equal(a.gt_inds, a0[0]) and torch.equal(a.max_overlaps, a0[1])


def test_rpn_proposals_stock_training_config_2000x5_levels():
    """The lineage's stock TRAINING proposal config: nms_pre = nms_post = max_num = 2000 on 5 levels - up to 10 000 kept
    candidates are merged per image (more than one 8192-entry sort).  Stage-wise bit-exact and end to end vs the C port."""
    cfgkw = dict(nms_pre=2000, nms_post=2000, max_num=2000, nms_thr=0.7)
    d = syn.cfg2(batch=2)
    props, nv, stages = _run_rpn(d, cfgkw)
    assert stages[3][:, :, 1].sum(axis=1).max() > 2000           # the merge really has to cut
    _check_rpn_stagewise(d, cfgkw, props, nv, stages)
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    ro, rn = cref.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfgkw)
    assert np.array_equal(nv, rn) and np.array_equal(props, ro) and np.all(nv == 2000)


# (1, 50, 68, 12000, 2000): the mx-rcnn TRAINING proposal config (rpn_pre_nms_top_n 12000 > the in-CTA sort capacity)
@pytest.mark.parametrize("N_,H,W,pre,post,scales", [(2, 38, 50, 6000, 300, (4, 8, 16, 32)), (1, 25, 34, 1000, 50, (8, 16, 32)),
                                                    (3, 7, 9, 200, 600, (2,)), (1, 50, 68, 12000, 2000, (4, 8, 16, 32))])
def test_multi_proposal_mx_layout(N_, H, W, pre, post, scales):
    """mx.nd.contrib.MultiProposal (NCHW in, cyclically padded rois out) vs the oracle: identical rows."""
    from mxdetection_b200.models.rpn_heads import MultiProposal
    rng = np.random.default_rng(N_ * 100 + H)
    A = 3 * len(scales)
    cls = (1.0 / (1.0 + np.exp(-rng.normal(-2, 2, (N_, 2 * A, H, W))))).astype(F)
    cls = np.round(cls * 64) / 64                                   # many exact score ties
    bbox = rng.normal(0, 0.4, (N_, 4 * A, H, W)).astype(F)
    info = np.array([[H * 16 - 5.0, W * 16 - 9.0, 1.5], [H * 16 - 40.0, W * 16 - 1.0, 1.0], [H * 12.0, W * 16.0, 0.7]], F)[:N_]
    kw = dict(rpn_pre_nms_top_n=pre, rpn_post_nms_top_n=post, threshold=0.7, rpn_min_size=16, scales=scales,
              ratios=(0.5, 1, 2), feature_stride=16)
    rois, sc = MultiProposal(T(cls.astype(F)), T(bbox), T(info), output_score=True, **kw)
    rref, sref = oracle.multi_proposal(cls.astype(F), bbox, info, **kw)
    assert np.array_equal(N(sc), sref)
    assert np.abs(N(rois) - rref).max() <= 1e-5 and np.array_equal(N(rois)[:, 0], rref[:, 0])


# ===================================================== SURVEY 8(f) next rows N1 / N2 / N4 ==
def test_random_sampler_and_target_packing_n1():
    from mxdetection_b200.core.bbox import MaxIoUAssigner, RandomSampler, pack_targets
    rng = np.random.default_rng(41)
    base = oracle.gen_base_anchors(8, [8], [0.5, 1, 2])
    anchors = oracle.grid_anchors(base, 40, 56, 8)
    gts = syn.gt_boxes(rng, 320, 448, 9)
    a = MaxIoUAssigner(0.7, 0.3, 0.3).assign(T(anchors), T(gts))
    ra, _, _ = oracle.max_iou_assign(anchors, gts, None, 0.7, 0.3, 0.3)
    keys = rng.random(anchors.shape[0]).astype(F)
    keys[::7] = keys[3]                                            # ties
    for num, frac, ub in ((256, 0.5, -1), (64, 0.25, 3), (8, 0.5, 0), (100, 0.7, -1), (200, 0.35, -1), (10, 0.9, 2)):   # 0.7 / 0.35 / 0.9 round down as fp32
        s = RandomSampler(num, frac, ub).sample(a.gt_inds, T(keys))
        pos, neg = oracle.targets.random_sample(ra, keys, num, frac, ub)
        gp = N(s.pos_inds); gn = N(s.neg_inds)
        assert int(s.num_pos.item()) == len(pos) and int(s.num_neg.item()) == len(neg)
        assert np.array_equal(gp[gp >= 0], pos) and np.array_equal(gn[gn >= 0], neg)
        lab, lw, tgt, tw = pack_targets(T(anchors), a.gt_inds, T(gts), s)
        rl, rlw, rt, rtw = oracle.targets.pack_targets(anchors, ra, gts, pos, neg)
        assert np.array_equal(N(lab), rl) and np.array_equal(N(lw), rlw) and np.array_equal(N(tw), rtw)
        assert np.abs(N(tgt) - rt).max() <= 1e-6


def test_multiclass_nms_and_det_bboxes_n2():
    from mxdetection_b200.models.bbox_heads import get_det_bboxes, multiclass_nms
    rng = np.random.default_rng(5)
    n, C = 300, 21
    rois = np.concatenate([np.zeros((n, 1)), syn.gt_boxes(rng, 600, 800, n)], 1).astype(F)
    logits = rng.normal(0, 2, (n, C)); logits[:, 0] += 2
    score = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(F)
    score = (np.round(score * 256) / 256).astype(F)               # exact ties across classes
    for pred_cols in (4, 4 * C):
        pred = rng.normal(0, 1.0, (n, pred_cols)).astype(F)
        dets, labels, num = get_det_bboxes(T(rois), T(score), T(pred), (600, 800), 1.0, 0.05, 0.5, 100)
        rd, rl = oracle.targets.get_det_bboxes(rois, score, pred, (600, 800), 1.0, 0.05, 0.5, 100)
        k = int(num.item())
        assert k == len(rl) and np.array_equal(N(labels)[:k], rl) and np.all(N(labels)[k:] == -1)
        assert np.abs(N(dets)[:k] - rd).max() <= 1e-5 and np.array_equal(N(dets)[:k, 4], rd[:, 4])
    # rescale (scale_factor != 1) in both branches: decoded boxes and rois-only (bbox_pred None)
    for pred_ in (pred, None):
        dets, labels, num = get_det_bboxes(T(rois), T(score), None if pred_ is None else T(pred_), (600, 800), 1.6, 0.05, 0.5, 100)
        rd, rl = oracle.targets.get_det_bboxes(rois, score, pred_, (600, 800), 1.6, 0.05, 0.5, 100)
        k = int(num.item())
        assert k == len(rl) and np.array_equal(N(labels)[:k], rl) and np.abs(N(dets)[:k] - rd).max() <= 1e-5
    d2, l2, n2 = multiclass_nms(T(rois[:, 1:]), T(score), 2.0, 0.5, 10)          # nothing passes the threshold
    assert int(n2.item()) == 0 and np.all(N(l2) == -1)


def test_mask_target_n4_bit_exact():
    """uint8 masks read directly by mxd_mask_target, Spec A in strict fp32: targets AND the fp32 RoIAlign values are
    bit-exact against the oracle (RoIs inside, across and outside the mask, degenerate, bad GT index)."""
    from mxdetection_b200.core.mask import mask_target
    rng = np.random.default_rng(8)
    G, H, W = 5, 96, 128
    masks = np.zeros((G, H, W), np.uint8)
    for g in range(G):
        y, x = np.ogrid[:H, :W]
        masks[g] = ((y - rng.uniform(20, 70)) ** 2 / rng.uniform(100, 900) + (x - rng.uniform(30, 100)) ** 2 / rng.uniform(100, 1600)) <= 1
    props = syn.gt_boxes(rng, H, W, 60)
    props[:4] = [[-30, -20, 40, 50], [100, 80, 140, 110], [50, 50, 50, 50], [10.25, 3.5, 90.75, 95.9]]
    inds = rng.integers(0, G, 60).astype(np.int32)
    inds[7] = -1; inds[8] = G                                        # out of range -> zeros
    for S, sr in ((28, 2), (14, 2), (28, -1), (7, 3)):
        got = N(mask_target(T(props), T(inds), T(masks), S, sample_ratio=sr))
        ref = oracle.targets.mask_target(props, inds, masks, S, sr)
        assert got.shape == (60, S, S) and np.array_equal(got, ref), (S, sr)
        raw = N(mask_target(T(props), T(inds), T(masks), S, sample_ratio=sr, binarize=False))
        data = masks.astype(F)[:, None]
        rois = np.concatenate([inds.astype(F)[:, None], props[:, :4]], 1)
        assert np.array_equal(raw, oracle.roi_align_forward(data, rois, (S, S), 1.0, sr)[:, 0])
    assert np.all(got[7] == 0) and np.all(got[8] == 0) and got.sum() > 0


def test_mask_paste_n4_bit_exact():
    """mxd_paste_masks (FCNMaskHead.get_seg_masks): class-indexed S x S probabilities -> image-size uint8 masks."""
    from mxdetection_b200.models.mask_heads import get_seg_masks
    from mxdetection_b200.core.mask import paste_masks
    rng = np.random.default_rng(21)
    n, C, S, H, W = 12, 5, 28, 150, 203
    pred = rng.uniform(0, 1, (n, C, S, S)).astype(F)
    pred[0, :, :, :] = 0.5                                          # exactly at the threshold: strict > keeps it out
    boxes = np.concatenate([syn.gt_boxes(rng, H, W, n), rng.uniform(0, 1, (n, 1)).astype(F)], 1).astype(F)
    boxes[1, :4] = [-20, -10, 60, 70]                               # crosses the image border
    boxes[2, :4] = [30, 40, 30, 40]                                 # one pixel
    boxes[3, :4] = [W - 10, H - 10, W + 30, H + 30]
    labels = rng.integers(0, C - 1, n).astype(np.int32)
    for scale in (1.0, 1.6):
        got = N(get_seg_masks(T(pred), T(boxes), T(labels), (H, W, 3), scale_factor=scale))
        ref = oracle.targets.paste_masks(pred, boxes, (H, W), labels, scale)
        assert got.shape == (n, H, W) and np.array_equal(got, ref), scale
    assert got[0].sum() == 0 and ref.sum() > 0
    single = N(paste_masks(T(pred[:, 1].copy()), T(boxes), (H, W)))
    assert np.array_equal(single, oracle.targets.paste_masks(pred[:, 1], boxes, (H, W)))


def test_empty_and_one_sided_inputs():
    """No RoIs at all, an image without RoIs, no positives / no negatives to sample, no GT, no candidate above the
    score threshold: shapes, zero fills and counts are what the reference's loops leave behind."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward, roi_align_backward
    from mxdetection_b200.core.bbox import MaxIoUAssigner, RandomSampler, pack_targets
    from mxdetection_b200.models.bbox_heads import get_det_bboxes
    rng = np.random.default_rng(12)
    shapes = [(2, 32, 40, 56), (2, 32, 20, 28)]
    feats = [T(rng.normal(0, 1, s).astype(F)) for s in shapes]
    scales = [0.25, 0.125]
    none = T(np.zeros((0, 5), F))
    assert tuple(roi_align_fpn_forward(feats, none, (7, 7), scales, 2).shape) == (0, 32, 7, 7)
    g = roi_align_fpn_backward(T(np.zeros((0, 32, 7, 7), F)), none, shapes, (7, 7), scales, 2,
                               grad_feats=[torch.full(s, 7.0, device="cuda") for s in shapes])
    assert all(float(x.abs().max()) == 0.0 for x in g)                 # req=write: zero-filled
    g1 = roi_align_backward(T(np.zeros((0, 32, 7, 7), F)), none, shapes[0], (7, 7), 0.25, 2)
    assert float(g1.abs().max()) == 0.0
    # RoIs on image 1 only: image 0's gradient planes are zeros, the output matches the oracle
    rois = np.concatenate([np.ones((30, 1)), syn.gt_boxes(rng, 160, 224, 30)], 1).astype(F)
    gout = rng.normal(0, 1, (30, 32, 7, 7)).astype(F)
    out = roi_align_fpn_forward(feats, T(rois), (7, 7), scales, 2)
    lv = oracle.map_roi_levels(rois, 2)
    assert close(N(out), cref.roi_align_forward([N(f) for f in feats], rois, (7, 7), scales, 2, lv), 1e-5)
    g = roi_align_fpn_backward(T(gout), T(rois), shapes, (7, 7), scales, 2)
    gref = cref.roi_align_backward(gout, rois, shapes, (7, 7), scales, 2, lv)
    assert all(float(x[0].abs().max()) == 0.0 for x in g) and all(close(N(x), r, 1e-4) for x, r in zip(g, gref))
    # sampler: only negatives / only positives / nothing assigned
    n = 5000
    keys = T(rng.random(n).astype(F))
    anchors = T(syn.gt_boxes(rng, 600, 800, n))
    gts = T(syn.gt_boxes(rng, 600, 800, 3))
    for assigned in (np.zeros(n, np.int32), np.ones(n, np.int32), np.full(n, -1, np.int32)):
        s_ = RandomSampler(64, 0.25, -1).sample(T(assigned), keys)
        pos, neg = oracle.targets.random_sample(assigned, N(keys), 64, 0.25, -1)
        assert int(s_.num_pos.item()) == len(pos) and int(s_.num_neg.item()) == len(neg)
        gp, gn = N(s_.pos_inds), N(s_.neg_inds)
        assert np.array_equal(gp[gp >= 0], pos) and np.array_equal(gn[gn >= 0], neg)
        lab, lw, tgt, tw = pack_targets(anchors, T(assigned), gts, s_)
        rl, rlw, rt, rtw = oracle.targets.pack_targets(N(anchors), assigned, N(gts), pos, neg)
        assert np.array_equal(N(lab), rl) and np.array_equal(N(lw), rlw) and np.array_equal(N(tw), rtw)
        assert np.abs(N(tgt) - rt).max() <= 1e-6
    # assigner without ground truth: everything negative (mmdet: assigned 0, overlaps 0)
    a0 = MaxIoUAssigner(0.5, 0.4, 0.2).assign_batch(anchors, T(np.zeros((1, 4, 4), F)), T(np.zeros(1, np.int32)), None)
    assert int(a0[0].abs().max()) == 0 and float(a0[1].abs().max()) == 0.0
    # detections: every score below the threshold -> num 0, rows 0 / -1
    rois_d = np.concatenate([np.zeros((50, 1)), syn.gt_boxes(rng, 600, 800, 50)], 1).astype(F)
    sc = np.full((50, 5), 0.01, F); sc[:, 0] = 0.96
    dets, labels, num = get_det_bboxes(T(rois_d), T(sc), T(rng.normal(0, 1, (50, 20)).astype(F)), (600, 800), 1.0, 0.05, 0.5, 100)
    assert int(num.item()) == 0 and float(dets.abs().max()) == 0.0 and np.all(N(labels) == -1)


def test_nms_threshold_extremes_and_identical_boxes():
    from mxdetection_b200.ops import nms_indices
    rng = np.random.default_rng(21)
    n = 700
    xy = rng.uniform(0, 400, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = rng.uniform(0, 1, n).astype(F)
    same = np.tile(boxes[:1], (n, 1))
    for bx, thr in ((boxes, 0.0), (boxes, 1.0), (boxes, 0.999999), (same, 0.5), (same, 1.0)):
        for delta in (0.0, 1.0):
            ref = oracle.nms(bx, scores, thr, delta=delta)
            keep, num = nms_indices(T(bx), T(scores), thr, delta=delta)
            assert int(num.item()) == len(ref) and np.array_equal(N(keep)[: len(ref)], ref), (thr, delta)


def test_roi_align_rois_outside_inverted_and_whole_image(roi_path):
    """RoIs that miss the map, inverted corners (width / height clamp to 1), RoIs larger than the map, single-pixel RoIs:
    the planners reject what does not fit their windows and the gather path takes over - same numbers either way."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    rng = np.random.default_rng(31)
    shapes = [(2, 32, 48, 64), (2, 32, 24, 32), (2, 32, 12, 16)]
    scales = [0.25, 0.125, 0.0625]
    feats = [rng.normal(0, 1, s).astype(F) for s in shapes]
    rois = np.array([[0, -500, -500, -300, -300], [0, 400, 300, 900, 700], [1, 100, 80, 60, 40], [1, -50, -50, 400, 300],
                     [0, 10.3, 20.7, 10.3, 20.7], [1, 0, 0, 255, 191], [0, 250, 5, 259, 190], [1, 3, 180, 250, 195],
                     [0, -20, 30, 15, 60], [1, 255.5, 191.5, 300, 200]], F)
    rois = np.concatenate([rois, np.concatenate([rng.integers(0, 2, (40, 1)), syn.gt_boxes(rng, 192, 256, 40)], 1).astype(F)])
    rois = rois[np.argsort(rois[:, 0], kind="stable")]
    lv = oracle.map_roi_levels(rois, 3)
    for ps in ((7, 7), (14, 14)):
        out = roi_align_fpn_forward([T(f) for f in feats], T(rois), ps, scales, 2)
        assert close(N(out), cref.roi_align_forward(feats, rois, ps, scales, 2, lv), 1e-5)
        gout = rng.normal(0, 1, (rois.shape[0], 32) + ps).astype(F)
        g = roi_align_fpn_backward(T(gout), T(rois), shapes, ps, scales, 2)
        gref = cref.roi_align_backward(gout, rois, shapes, ps, scales, 2, lv)
        assert all(close(N(x), r, 1e-4) for x, r in zip(g, gref))


def test_pipeline_is_cuda_graph_capturable():
    """No allocation / sync inside the library: the whole proposal stage replays from a CUDA graph."""
    from mxdetection_b200.models.rpn_heads import RPNHead, ProposalConfig
    d = syn.rpn_inputs(2, 2, 160, 224)
    head, cfg = RPNHead(), ProposalConfig(nms_pre=300, nms_post=100, max_num=150)
    sc = [T(s) for s in d["scores"]]; dl = [T(x) for x in d["deltas"]]; shp = T(d["img_shapes"])
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eager, _ = head.get_proposals(sc, dl, d["feat_shapes"], shp, cfg)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            props, nv = head.get_proposals(sc, dl, d["feat_shapes"], shp, cfg)
        props.zero_()
        g.replay()
        st.synchronize()
    assert torch.equal(props, eager)


def test_roi_stage_writes_exactly_its_outputs():
    """Output-side canaries: `out` and every gradient map are views into larger buffers whose guard zones (64 KiB either
    side, a NaN pattern) must survive the row-ring forward and the tile backward, and every byte in between is written."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=4, with_features=False)
    shapes = [(4, 256, h, w) for h, w in d["feat_shapes"]]
    gen = torch.Generator(device="cuda").manual_seed(11)
    feats = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    rois = T(d["rois"])
    gout = torch.randn((rois.shape[0], 256, 7, 7), device="cuda", generator=gen)
    G = 16384                                                  # guard floats

    def guarded(shape):
        n = int(np.prod(shape))
        full = torch.full((n + 2 * G,), float("nan"), device="cuda")
        return full, full[G:G + n].view(shape)

    f_out, out = guarded(tuple(gout.shape))
    gs = [guarded(s) for s in shapes]
    roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
    roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=[v for _, v in gs])
    torch.cuda.synchronize()
    for full, view in [(f_out, out)] + gs:
        assert bool(torch.isnan(full[:G]).all()) and bool(torch.isnan(full[-G:]).all()), "wrote outside the output tensor"
        assert not bool(torch.isnan(view).any()), "left output bytes unwritten"


def test_roi_align_many_images_many_rois(roi_path):
    """Unit / group / tile bookkeeping far from the benchmark's shape: 24 images, 6000 RoIs in RANDOM batch order (a
    tile's RoI list then spans the whole RoI array: several bitmap windows in the list sort), 3 levels, 32 channels."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    rng = np.random.default_rng(77)
    Nn, C, R = 24, 32, 6000
    shapes = [(Nn, C, 64, 96), (Nn, C, 32, 48), (Nn, C, 16, 24)]
    scales = [0.25, 0.125, 0.0625]
    feats = [rng.standard_normal(sh).astype(F) for sh in shapes]
    rois = np.concatenate([rng.integers(0, Nn, (R, 1)), syn.gt_boxes(rng, 256, 384, R)], 1).astype(F)
    lv = oracle.map_roi_levels(rois, 3)
    out = roi_align_fpn_forward([T(f) for f in feats], T(rois), (7, 7), scales, 2)
    assert close(N(out), cref.roi_align_forward(feats, rois, (7, 7), scales, 2, lv), 1e-5)
    gout = rng.standard_normal((R, C, 7, 7)).astype(F)
    g = roi_align_fpn_backward(T(gout), T(rois), shapes, (7, 7), scales, 2)
    gref = cref.roi_align_backward(gout, rois, shapes, (7, 7), scales, 2, lv)
    assert all(close(N(a), b, 1e-4) for a, b in zip(g, gref))
    g2 = roi_align_fpn_backward(T(gout), T(rois), shapes, (7, 7), scales, 2)
    if roi_path != "gather":
        assert all(torch.equal(a, b) for a, b in zip(g, g2))        # list order fixed -> bit-reproducible


def test_roi_stage_is_cuda_graph_capturable():
    """The FPN RoI stage of the benchmarked kind (row-ring forward with tensor-map TMA, tile backward, programmatic
    dependent launches between planners and main kernels) captures into a CUDA graph and replays bit-identically."""
    from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
    d = syn.cfg3(batch=4, with_features=False)          # 4 images: enough units for the ring forward
    shapes = [(4, 256, h, w) for h, w in d["feat_shapes"]]
    gen = torch.Generator(device="cuda").manual_seed(9)
    feats = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    rois = T(d["rois"])
    gout = torch.randn((rois.shape[0], 256, 7, 7), device="cuda", generator=gen)
    out = torch.empty_like(gout)
    grads = [torch.empty(s, device="cuda") for s in shapes]
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
        roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)
        st.synchronize()
        eager_out, eager_g = out.clone(), [t.clone() for t in grads]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
            roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)
        for _ in range(3):
            out.fill_(float("nan"))
            for t in grads:
                t.fill_(float("nan"))
            g.replay()
            st.synchronize()
            assert torch.equal(out, eager_out) and all(torch.equal(a, b) for a, b in zip(grads, eager_g))


# ============================================================ SURVEY 8(e): tail gather ==
def test_pack_detections_layout():
    from mxdetection_b200.parallel import pack_detections, unpack_detections
    from test_parallel_gloo import pack_reference
    rng = np.random.default_rng(3)
    props = rng.standard_normal((5, 7, 5)).astype(F)
    nv = np.array([7, 0, 3, 1, 6], np.int32)
    packed = pack_detections(T(props), T(nv), 40)
    assert np.array_equal(N(packed), pack_reference(torch.from_numpy(props), torch.from_numpy(nv), 40).numpy())
    dets, counts, ids = unpack_detections(packed)
    assert N(counts).tolist() == nv.tolist() and N(ids).tolist() == [40, 41, 42, 43, 44]


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from mxdetection_b200.models.rpn_heads import ProposalConfig, RPNHead
        from mxdetection_b200.parallel import gather_detections, shard_range
        n_img = 2 * world
        d = syn.rpn_inputs(11, n_img, 160, 224)
        lo, hi = shard_range(n_img, rank, world)
        dev = "cuda:%d" % rank
        sc = [torch.from_numpy(s[lo:hi]).to(dev) for s in d["scores"]]
        dl = [torch.from_numpy(x[lo:hi]).to(dev) for x in d["deltas"]]
        shp = torch.from_numpy(d["img_shapes"][lo:hi]).to(dev)
        props, nv = RPNHead().get_proposals(sc, dl, d["feat_shapes"], shp, ProposalConfig(nms_pre=300, nms_post=100, max_num=150, nms_thr=0.7))
        dets, counts = gather_detections(props, nv, first_image_id=lo)
        torch.cuda.synchronize()
        q.put((rank, dets.cpu().numpy(), counts.cpu().numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        q.put((rank, repr(e), None))


def test_nccl_gather_equals_single_gpu_concatenation():
    """the detections gathered over NCCL from W GPUs (2 images each) equal, bit for bit, the proposals of
    one GPU run over all 2W images.  Needs >= 2 GPUs (gpurun --gpus N); on a one-GPU box it is skipped."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from mxdetection_b200.models.rpn_heads import ProposalConfig, RPNHead
    from mxdetection_b200.parallel import pack_detections, unpack_detections
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n_img = 2 * world
    d = syn.rpn_inputs(11, n_img, 160, 224)
    props, nv = RPNHead().get_proposals([T(s) for s in d["scores"]], [T(x) for x in d["deltas"]], d["feat_shapes"], T(d["img_shapes"]),
                                        ProposalConfig(nms_pre=300, nms_post=100, max_num=150, nms_thr=0.7))
    ref_d, ref_c, _ = unpack_detections(pack_detections(props, nv, 0))
    for rank, dets, counts in res:
        assert counts is not None, dets
        assert np.array_equal(dets, N(ref_d)) and np.array_equal(counts, N(ref_c)), "rank %d" % rank
