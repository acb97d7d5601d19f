"""GPU parity tests of the HBM-bound kernels (K2-K7) against the CPU oracle / torch fp32 references."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def test_frames_to_f16c4(lib):
    from vehicle_counting_b200 import ops
    rng = np.random.default_rng(0)
    for shape in [(2, 64, 48), (1, 7, 9), (3, 640, 640)]:
        fr = torch.from_numpy(rng.integers(0, 256, shape + (3,), dtype=np.uint8))
        out = torch.full(shape + (4,), 5.0, dtype=torch.float16, device=DEV)
        ops.frames_to_f16c4(fr.to(DEV), out)
        ref = (fr.float() / 255.0).half()
        got = out.cpu()
        assert torch.equal(got[..., :3], ref)          # bit-exact: same fp32 divide, same rounding
        assert (got[..., 3] == 0).all()


def test_frames_to_f16_s2d(lib):
    from vehicle_counting_b200 import ops
    rng = np.random.default_rng(1)
    for shape in [(2, 64, 48), (1, 2, 2), (3, 640, 640)]:
        fr = torch.from_numpy(rng.integers(0, 256, shape + (3,), dtype=np.uint8))
        n, h, w = shape
        out = torch.full((n, h // 2, w // 2, 16), 5.0, dtype=torch.float16, device=DEV)
        ops.frames_to_f16_s2d(fr.to(DEV), out)
        x = (fr.float() / 255.0).half()                                   # [n, h, w, 3]
        ref = x.view(n, h // 2, 2, w // 2, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 12)
        got = out.cpu()
        assert torch.equal(got[..., :12], ref)                            # bit-exact
        assert (got[..., 12:] == 0).all()


def test_frames_to_f16_s2d_wpad(lib):
    """the W-padded ingest of the row-window stem: same pixels at column x + 1, pad columns and the tail untouched"""
    from vehicle_counting_b200 import ops
    rng = np.random.default_rng(2)
    for shape in [(2, 64, 48), (1, 2, 2), (2, 6, 10), (2, 384, 640)]:       # widths with and without the 8-pixel fast path
        fr = torch.from_numpy(rng.integers(0, 256, shape + (3,), dtype=np.uint8))
        n, h, w = shape
        h2, w2 = h // 2, w // 2
        flat = torch.full((n * h2 * (w2 + 2) * 16 + 16,), 5.0, dtype=torch.float16, device=DEV)
        ops.frames_to_f16_s2d_wpad(fr.to(DEV), flat)
        x = (fr.float() / 255.0).half()
        ref = x.view(n, h2, 2, w2, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(n, h2, w2, 12)
        got = flat.cpu()
        body = got[:n * h2 * (w2 + 2) * 16].view(n, h2, w2 + 2, 16)
        assert torch.equal(body[:, :, 1:w2 + 1, :12], ref)
        assert (body[:, :, 1:w2 + 1, 12:] == 0).all()
        assert (body[:, :, 0] == 5.0).all() and (body[:, :, w2 + 1] == 5.0).all() and (got[-16:] == 5.0).all()


def test_upsample2x_into_slice(lib):
    from vehicle_counting_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, h, w, c = 2, 5, 7, 16
    src = torch.randn(n, h, w, 24, generator=g).half()
    dst = torch.full((n, 2 * h, 2 * w, 40), 9.0, dtype=torch.float16)
    sd, dd = src.to(DEV), dst.to(DEV)
    ops.upsample2x(sd[..., 8:], 24, dd[..., 16:], 40, n, h, w, c)
    ref = F.interpolate(src[..., 8:24].permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).half()
    got = dd.cpu()
    assert torch.equal(got[..., 16:32], ref)
    assert (got[..., :16] == 9.0).all() and (got[..., 32:] == 9.0).all()


@pytest.mark.parametrize("h,w", [(20, 20), (12, 23), (40, 40)])
def test_sppf_pool(lib, h, w):
    from vehicle_counting_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, c = 3, 32
    buf = torch.zeros(n, h, w, 4 * c, dtype=torch.float16)
    buf[..., :c] = torch.randn(n, h, w, c, generator=g).half()
    bd = buf.to(DEV)
    ops.sppf_pool(bd, 4 * c, n, h, w, c)
    x = buf[..., :c].permute(0, 3, 1, 2).float()
    y1 = F.max_pool2d(x, 5, 1, 2); y2 = F.max_pool2d(y1, 5, 1, 2); y3 = F.max_pool2d(y2, 5, 1, 2)
    ref = torch.cat([x, y1, y2, y3], 1).permute(0, 2, 3, 1).half()
    assert torch.equal(bd.cpu(), ref)


def test_maxpool_reid_stem(lib):
    from vehicle_counting_b200 import ops
    g = torch.Generator().manual_seed(2)
    n, h, w, c = 4, 50, 50, 64
    x = torch.randn(n, h, w, c, generator=g).half()
    out = torch.empty(n, 25, 25, c, dtype=torch.float16, device=DEV)
    ops.maxpool(x.to(DEV), c, out, c, n, h, w, c, 3, 2, 1)
    ref = F.max_pool2d(x.permute(0, 3, 1, 2).float(), 3, 2, 1).permute(0, 2, 3, 1).half()
    assert torch.equal(out.cpu(), ref)


def test_avgpool_l2norm(lib):
    from vehicle_counting_b200 import ops
    g = torch.Generator().manual_seed(3)
    n, hw, c = 9, 16, 512
    x = torch.randn(n, hw, c, generator=g).abs().half()
    out = torch.empty(n, c, dtype=torch.float32, device=DEV)
    ops.avgpool_l2norm(x.to(DEV), c, n, hw, c, out)
    m = x.float().mean(1)
    ref = m / m.norm(p=2, dim=1, keepdim=True)
    assert (out.cpu() - ref).abs().max().item() < 2e-6


def test_bn_train_stats_and_apply(lib):
    from vehicle_counting_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(4)
    c, hw = 64, 25
    crops = [3, 1, 5]
    rows = sum(crops) * hw
    x = torch.randn(rows, c, generator=g) * 2 + 0.5
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g)
    seg_start = torch.tensor(np.cumsum([0] + [k * hw for k in crops]), dtype=torch.int32)
    row_seg = torch.repeat_interleave(torch.arange(3, dtype=torch.int32), torch.tensor([k * hw for k in crops]))
    res = torch.randn(rows, c, generator=g).half()
    xd = x.to(DEV)
    scale = torch.empty(3, c, device=DEV); shift = torch.empty(3, c, device=DEV)
    ops.bn_train_stats(xd, c, seg_start.to(DEV), 3, gamma.to(DEV), beta.to(DEV), 1e-5, scale, shift)
    y = torch.empty(rows, c, dtype=torch.float16, device=DEV)
    ops.bn_apply(xd, c, rows, row_seg.to(DEV), scale, shift, res.to(DEV), c, L.ACT_RELU, y, c)
    refs = []
    for s in range(3):
        xs = x[seg_start[s]:seg_start[s + 1]]
        # [rows, c] -> BatchNorm over rows == BatchNorm2d over (N,H,W) in train mode
        refs.append(F.batch_norm(xs.t().unsqueeze(0), None, None, gamma, beta, True, 0.0, 1e-5).squeeze(0).t())
    ref = F.relu(torch.cat(refs, 0) + res.float())
    assert (y.float().cpu() - ref).abs().max().item() < 4e-3 * ref.abs().max().item()


def _decode_inputs(n, levels, nc, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    no = nc + 5
    logits = []
    for (ny, nx) in levels:
        t = torch.randn(n, ny, nx, 256 if 3 * no <= 256 else 3 * no, generator=g) * 2.0
        t[..., 4::no][..., :3] -= 1.0
        logits.append(t.to(dtype))
    return logits


def _oracle_pred(logits, levels, nc):
    """[n, P, no] decoded predictions exactly like oracle.yolov5.Detect.forward (fp32 CPU)."""
    from oracle import yolov5 as Y
    no = nc + 5
    z = []
    for i, (ny, nx) in enumerate(levels):
        x = logits[i].float()[..., :3 * no]
        n = x.shape[0]
        x = x.view(n, ny, nx, 3, no).permute(0, 3, 1, 2, 4)          # [n, a, y, x, no]
        yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
        anchor = torch.tensor(Y.ANCHORS_PX[i], dtype=torch.float32).view(1, 3, 1, 1, 2)
        y = x.sigmoid()
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * Y.STRIDES[i]
        wh = (y[..., 2:4] * 2.0) ** 2 * anchor
        z.append(torch.cat((xy, wh, y[..., 4:]), -1).reshape(n, -1, no))
    return torch.cat(z, 1)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_detect_decode_and_nms_match_oracle(lib, dtype):
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import ops, _lib as L
    n, nc = 3, 80
    levels = [(16, 20), (8, 10), (4, 5)]
    logits = _decode_inputs(n, levels, nc, 7, dtype)
    pred = _oracle_pred(logits, levels, nc)
    P = pred.shape[1]
    dd = L.DetectDesc()
    dd.n, dd.nc, dd.num_levels = n, nc, 3
    dd.logits_dtype = L.F32 if dtype == torch.float32 else L.F16
    dd.conf_thres, dd.max_candidates = 0.25, P
    dl = [t.to(DEV) for t in logits]
    for i, (ny, nx) in enumerate(levels):
        lv = dd.level[i]
        lv.logits, lv.pitch, lv.ny, lv.nx, lv.stride = dl[i].data_ptr(), dl[i].shape[-1], ny, nx, float(Y.STRIDES[i])
        for a in range(3):
            lv.anchor_w[a] = float(Y.ANCHORS_PX[i][2 * a]); lv.anchor_h[a] = float(Y.ANCHORS_PX[i][2 * a + 1])
    cb = torch.zeros(n, P, 4, device=DEV); cs = torch.zeros(n, P, device=DEV)
    cc = torch.zeros(n, P, dtype=torch.int32, device=DEV); ci = torch.zeros(n, P, dtype=torch.int32, device=DEV)
    cnt = torch.zeros(n, dtype=torch.int32, device=DEV)
    ops.detect_decode(dd, cb, cs, cc, ci, cnt)
    nd = L.NmsDesc()
    nd.n, nd.max_candidates, nd.max_det, nd.iou_thres, nd.max_wh, nd.max_nms = n, P, 300, 0.45, 4096.0, 30000
    ws = torch.zeros(ops.nms_workspace_bytes(n, P) // 8, dtype=torch.int64, device=DEV)
    det = torch.zeros(n, 300, 6, device=DEV); dc = torch.zeros(n, dtype=torch.int32, device=DEV)
    ops.nms(nd, cb, cs, cc, ci, cnt, ws, det, dc)
    torch.cuda.synchronize()
    ref = Y.non_max_suppression(pred, 0.25, 0.45, None, 300, 4096)
    cnt_h, dc_h, det_h = cnt.cpu(), dc.cpu(), det.cpu()
    for b in range(n):
        # candidate set: same prediction indices as the oracle's filter (threshold band excluded)
        x = pred[b]
        conf = (x[:, 5:] * x[:, 4:5]).max(1).values
        keep = (x[:, 4] > 0.25) & (conf > 0.25)
        band = ((x[:, 4] - 0.25).abs() < 1e-4) | ((conf - 0.25).abs() < 1e-4)
        got_idx = set(ci[b, :cnt_h[b]].cpu().tolist())
        want = set(torch.nonzero(keep & ~band).flatten().tolist())
        maybe = set(torch.nonzero(band).flatten().tolist())
        assert want <= got_idx <= (want | maybe)
        r = ref[b]
        assert dc_h[b].item() == r.shape[0], (b, dc_h[b].item(), r.shape[0])
        got = det_h[b, :r.shape[0]]
        assert torch.equal(got[:, 5], r[:, 5])
        assert (got[:, :4] - r[:, :4]).abs().max().item() < 1e-2
        assert (got[:, 4] - r[:, 4]).abs().max().item() < 1e-5
        assert (got[1:, 4] <= got[:-1, 4]).all()


def test_nms_bit_exact_on_given_candidates(lib):
    """Integer/index work is bit-exact: same candidates in -> same kept indices, order and count out."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import ops, _lib as L
    rng = np.random.default_rng(5)
    n = 4
    counts = [0, 1, 700, 9000]            # empty, single, typical, > smem sort capacity (global path)
    P = 9000
    cb = np.zeros((n, P, 4), np.float32); cs = np.zeros((n, P), np.float32)
    cc = np.zeros((n, P), np.int32); ci = np.zeros((n, P), np.int32)
    for b, c in enumerate(counts):
        ctr = rng.uniform(0, 640, (c, 2)).astype(np.float32)
        wh = rng.uniform(8, 120, (c, 2)).astype(np.float32)
        cb[b, :c, :2] = ctr - wh / 2; cb[b, :c, 2:] = ctr + wh / 2
        s = rng.uniform(0.25, 1.0, c).astype(np.float32)
        s[: c // 3] = np.round(s[: c // 3], 2)          # force score ties
        cs[b, :c] = s
        cc[b, :c] = rng.integers(0, 4, c)
        ci[b, :c] = rng.permutation(c * 3)[:c]          # unique, unordered prediction indices
    nd = L.NmsDesc()
    nd.n, nd.max_candidates, nd.max_det, nd.iou_thres, nd.max_wh, nd.max_nms = n, P, 300, 0.45, 4096.0, 30000
    t = lambda a: torch.from_numpy(a).to(DEV)
    ws = torch.zeros(ops.nms_workspace_bytes(n, P) // 8, dtype=torch.int64, device=DEV)
    det = torch.zeros(n, 300, 6, device=DEV); dc = torch.zeros(n, dtype=torch.int32, device=DEV)
    ops.nms(nd, t(cb), t(cs), t(cc), t(ci), t(np.asarray(counts, np.int32)), ws, det, dc)
    torch.cuda.synchronize()
    det_h, dc_h = det.cpu().numpy(), dc.cpu().numpy()
    for b, c in enumerate(counts):
        order = np.lexsort((ci[b, :c], -cs[b, :c]))      # score desc, prediction index asc
        boxes = cb[b, :c][order] + (cc[b, :c][order].astype(np.float32) * np.float32(4096.0))[:, None]
        keep = Y.greedy_nms(boxes, np.arange(c, 0, -1, dtype=np.float32), 0.45)[:300]
        sel = order[keep]
        assert dc_h[b] == len(sel)
        np.testing.assert_array_equal(det_h[b, :len(sel), :4], cb[b, sel])
        np.testing.assert_array_equal(det_h[b, :len(sel), 4], cs[b, sel])
        np.testing.assert_array_equal(det_h[b, :len(sel), 5], cc[b, sel].astype(np.float32))


def test_roi_resize_norm_matches_oracle(lib):
    from oracle import reid as R
    from vehicle_counting_b200 import ops, _lib as L
    rng = np.random.default_rng(11)
    fh, fw = 360, 480
    frames = rng.integers(0, 256, (2, fh, fw, 3), dtype=np.uint8)
    boxes = np.array([[10.2, 20.7, 90.9, 200.1], [-5.0, -3.0, 40.0, 50.0], [400.5, 300.2, 600.0, 400.0],
                      [100.0, 100.0, 150.0, 150.0], [30.3, 40.4, 33.9, 45.1], [0.0, 0.0, 479.9, 359.9]], np.float64)
    frame_of = np.array([0, 1, 0, 1, 0, 1], np.int32)
    rois = torch.zeros(len(boxes), 5, dtype=torch.int32, device=DEV)
    ops.boxes_to_rois(torch.from_numpy(boxes).to(DEV), torch.from_numpy(frame_of).to(DEV), len(boxes), fw, fh, rois)
    want = np.array([(f,) + R.crop_box(b, fw, fh) for f, b in zip(frame_of, boxes)], np.int32)
    np.testing.assert_array_equal(rois.cpu().numpy(), want)          # integer crop rule: bit-exact
    crops = [frames[f][y1:y2, x1:x2] for f, x1, y1, x2, y2 in want]
    ref = R.preprocess(crops).permute(0, 2, 3, 1)                   # NHWC fp32
    for oc in (4, 16):
        rd = L.RoiDesc()
        rd.num_rois, rd.out_size, rd.out_channels = len(boxes), 50, oc
        for c in range(3):
            rd.mean[c] = R.NORM_MEAN[c]; rd.inv_std[c] = 1.0 / R.NORM_STD[c]
        out = torch.full((len(boxes), 50, 50, oc), 3.0, dtype=torch.float16, device=DEV)
        ops.roi_resize_norm(rd, torch.from_numpy(frames).to(DEV), fh, fw, rois, out)
        got = out.float().cpu()
        assert (got[..., 3:] == 0).all()
        assert (got[..., :3] - ref).abs().max().item() < 2.5e-3    # fp16 storage of values up to ~2.7


def test_reid_fused_stem_matches_unfused_reference(lib):
    """roi_stem_patches + reid_stem_pool (crop -> conv3x3 3->64 + bias -> ReLU -> maxpool 3/2/1 in one pass, tcgen05) against
    the same arithmetic spelled out with torch on the fp16 crop that vcb_roi_resize_norm produces
    (/root/reference/networks/deepsort/deep/model.py:52-60)."""
    import torch.nn.functional as F
    from oracle import reid as R
    from vehicle_counting_b200 import ops, _lib as L
    rng = np.random.default_rng(5)
    fh, fw = 240, 320
    frames = torch.from_numpy(rng.integers(0, 256, (2, fh, fw, 3), dtype=np.uint8)).to(DEV)
    n = 37                                                     # not a multiple of anything: exercises the tile tail
    wh = rng.uniform(8, 200, (n, 2)); tl = rng.uniform(0, 1, (n, 2)) * (np.array([fw, fh]) - wh)
    rois_np = np.concatenate([rng.integers(0, 2, (n, 1)), tl, tl + wh], 1).astype(np.int32)
    rois_np[5] = (0, 10, 10, 10, 30)                           # zero-width crop: the unfused kernel writes zeros, so must this one
    rois = torch.from_numpy(rois_np).to(DEV)
    rd = L.RoiDesc()
    rd.num_rois, rd.out_size, rd.out_channels = n, 50, 4
    for c in range(3):
        rd.mean[c] = R.NORM_MEAN[c]; rd.inv_std[c] = 1.0 / R.NORM_STD[c]
    x = torch.zeros(n, 50, 50, 4, dtype=torch.float16, device=DEV)
    ops.roi_resize_norm(rd, frames, fh, fw, rois, x)
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.3)
    b = torch.randn(64, generator=g) * 0.2
    wp, bp = ops.pack_reid_stem_weights(w.to(DEV), b.to(DEV))
    patches = torch.full((n, 25, 128, 32), 7.0, dtype=torch.float16, device=DEV)
    out = torch.full((n, 25, 25, 64), -5.0, dtype=torch.float16, device=DEV)
    ops.roi_stem_patches(rd, frames, fh, fw, rois, patches)
    ops.reid_stem_pool(patches, wp, bp, out, n)
    torch.cuda.synchronize()
    assert tuple(lib_fault()) == (0, 0, 0, 0)
    # the im2col operand itself is integer-exact: element (r*3+s)*3+c of row ti*11+tj of block (by,bx) = crop[y+r-1, x+s-1, c]
    xc = F.pad(x[..., :3].float().cpu().permute(0, 3, 1, 2), (1, 1, 1, 1))                     # [n,3,52,52]
    cols = F.unfold(xc, 3).view(n, 3, 9, 50, 50).permute(0, 3, 4, 2, 1).reshape(n, 50, 50, 27)   # k = tap*3 + c
    pt = patches.float().cpu()
    for blk in (0, 7, 24):
        by, bx = divmod(blk, 5)
        for row in (0, 12, 60, 120):
            ti, tj = divmod(row, 11)
            cy, cx = 10 * by - 1 + ti, 10 * bx - 1 + tj
            want = cols[:, cy, cx] if (cy >= 0 and cx >= 0) else torch.zeros(n, 27)
            assert torch.equal(pt[:, blk, row, :27], want), (blk, row)
    assert (pt[:, :, 121:, :] == 0).all() and (pt[..., 27:] == 0).all()
    # conv + bias + ReLU (fp32 accumulate on the fp16 operands), fp16 rounding, max-pool
    conv = F.conv2d(x[..., :3].float().cpu().permute(0, 3, 1, 2), wp[:, :27].float().cpu().view(64, 3, 3, 3).permute(0, 3, 1, 2), b, 1, 1)
    ref = F.max_pool2d(F.relu(conv).half().float(), 3, 2, 1).permute(0, 2, 3, 1)
    got = out.float().cpu()
    assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_reid_fused_stem_train_mode_bn_matches_torch(lib):
    """The fused stem under train-mode BatchNorm (model.py:52-60 with the reference's batch statistics, one segment per
    Extractor call): vcb_reid_stem_stats + vcb_bn_seg_finalize + vcb_reid_stem_pool_bn against F.batch_norm(training=True) per
    segment on the same fp16 crop."""
    import torch.nn.functional as F
    from oracle import reid as R
    from vehicle_counting_b200 import ops, _lib as L
    rng = np.random.default_rng(6)
    fh, fw = 240, 320
    frames = torch.from_numpy(rng.integers(0, 256, (2, fh, fw, 3), dtype=np.uint8)).to(DEV)
    n, nb = 37, 64                                             # 37 real crops in a bucket of 64: padding crops form their own segment
    seg_sizes = [5, 1, 20, 11]
    wh = rng.uniform(8, 200, (nb, 2)); tl = rng.uniform(0, 1, (nb, 2)) * (np.array([fw, fh]) - wh)
    rois_np = np.concatenate([rng.integers(0, 2, (nb, 1)), tl, tl + wh], 1).astype(np.int32)
    rois_np[n:] = 0
    rois = torch.from_numpy(rois_np).to(DEV)
    rd = L.RoiDesc()
    rd.num_rois, rd.out_size, rd.out_channels = nb, 50, 4
    for c in range(3):
        rd.mean[c] = R.NORM_MEAN[c]; rd.inv_std[c] = 1.0 / R.NORM_STD[c]
    x = torch.zeros(nb, 50, 50, 4, dtype=torch.float16, device=DEV)
    ops.roi_resize_norm(rd, frames, fh, fw, rois, x)
    g = torch.Generator().manual_seed(4)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(64, generator=g) * 0.2
    gamma = torch.rand(64, generator=g) + 0.5
    beta = torch.randn(64, generator=g) * 0.1
    wp, bp = ops.pack_reid_stem_weights(w.to(DEV), b.to(DEV))
    nseg = len(seg_sizes)
    soc = np.full(nb, nseg, np.int32); soc[:n] = np.repeat(np.arange(nseg), seg_sizes)
    cnt = np.zeros(nseg + 1, np.int32); cnt[:nseg] = seg_sizes; cnt[nseg] = nb - n
    soc_d, cnt_d = torch.from_numpy(soc).to(DEV), torch.from_numpy(cnt).to(DEV)
    sums = torch.zeros(nseg + 1, 64, 2, dtype=torch.float64, device=DEV)
    affine = torch.zeros(nseg + 1, 64, 2, dtype=torch.float32, device=DEV)
    patches = torch.zeros(nb, 25, 128, 32, dtype=torch.float16, device=DEV)
    out = torch.full((nb, 25, 25, 64), -5.0, dtype=torch.float16, device=DEV)
    ops.roi_stem_patches(rd, frames, fh, fw, rois, patches)
    ops.reid_stem_stats(patches, wp, bp, nb, soc_d, sums)
    ops.bn_seg_finalize(sums, cnt_d, nseg + 1, 64, 2500, gamma.to(DEV), beta.to(DEV), bp, 1e-5, affine)
    ops.reid_stem_pool_bn(patches, wp, affine, soc_d, out, nb)
    torch.cuda.synchronize()
    assert tuple(lib_fault()) == (0, 0, 0, 0)
    conv = F.conv2d(x[:n, ..., :3].float().cpu().permute(0, 3, 1, 2), wp[:, :27].float().cpu().view(64, 3, 3, 3).permute(0, 3, 1, 2), b, 1, 1)
    got_sums = sums.cpu()
    off = 0
    for si, k in enumerate(seg_sizes):
        c_ = conv[off:off + k].double()
        np.testing.assert_allclose(got_sums[si, :, 0].numpy(), c_.sum((0, 2, 3)).numpy(), rtol=2e-4, atol=2e-2)
        np.testing.assert_allclose(got_sums[si, :, 1].numpy(), (c_ * c_).sum((0, 2, 3)).numpy(), rtol=2e-4, atol=2e-2)
        y = F.relu(F.batch_norm(conv[off:off + k], None, None, gamma, beta, True, 0.0, 1e-5))
        ref = F.max_pool2d(y.half().float(), 3, 2, 1).permute(0, 2, 3, 1)
        err = (out[off:off + k].float().cpu() - ref).abs().max().item()
        assert err <= 4e-3 * max(1.0, ref.abs().max().item()), (si, k, err)
        off += k


def _stem_case(seed, n, fh=240, fw=320, frames_n=2):
    from oracle import reid as R
    from vehicle_counting_b200 import ops, _lib as L
    rng = np.random.default_rng(seed)
    frames = torch.from_numpy(rng.integers(0, 256, (frames_n, fh, fw, 3), dtype=np.uint8)).to(DEV)
    wh = rng.uniform(8, 200, (n, 2)); tl = rng.uniform(0, 1, (n, 2)) * (np.array([fw, fh]) - wh)
    rois_np = np.concatenate([rng.integers(0, frames_n, (n, 1)), tl, tl + wh], 1).astype(np.int32)
    rd = L.RoiDesc()
    rd.num_rois, rd.out_size, rd.out_channels, rd.num_frames = n, 50, 4, frames_n
    for c in range(3):
        rd.mean[c] = R.NORM_MEAN[c]; rd.inv_std[c] = 1.0 / R.NORM_STD[c]
    return rng, frames, rois_np, rd


@pytest.mark.parametrize("n", [1, 37, 700])
def test_reid_direct_stem_matches_torch_on_the_resized_crop(lib, n):
    """vcb_reid_stem_direct (frames + ROIs -> crop -> conv3x3 3->64 + bias -> ReLU -> maxpool 3/2/1 in ONE kernel, the im2col rows
    built in shared memory) against torch on the fp16 crop vcb_roi_resize_norm produces, and against the two-kernel form that
    stages patches in HBM (/root/reference/networks/deepsort/deep/model.py:52-60).  n = 1 (one CTA), 37 (one crop per CTA, a
    degenerate ROI, a ROI of another frame index out of range) and 700 (several crops per CTA: the double-buffered crop)."""
    import torch.nn.functional as F
    from vehicle_counting_b200 import ops
    rng, frames, rois_np, rd = _stem_case(50 + n, n)
    fh, fw = 240, 320
    if n >= 37:
        rois_np[5] = (0, 10, 10, 10, 30)                         # zero-width crop -> zeros in, bias out
        rois_np[9, 0] = 7                                        # frame index out of range -> zeros in
        rois_np[11] = (1, 0, 0, 1, 1)                            # one pixel in the first corner (every tap clamps)
        rois_np[12] = (1, fw - 1, fh - 1, fw, fh)                # one pixel in the last corner: the last bytes of the last frame
        rois_np[13] = (0, 0, 0, fw, fh)                          # the whole frame (6.4x / 4.8x reduction)
        rois_np[14] = (1, 3, 5, 5, 6)                            # 2 x 1 pixels
        rois_np[15] = (0, fw - 2, 0, fw, fh)                     # two columns, full height
    rois = torch.from_numpy(rois_np).to(DEV)
    x = torch.zeros(n, 50, 50, 4, dtype=torch.float16, device=DEV)
    ops.roi_resize_norm(rd, frames, fh, fw, rois, x)
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.3)
    b = torch.randn(64, generator=g) * 0.2
    wpd = ops.pack_reid_stem_weights_direct(w.to(DEV), b.to(DEV))
    out = torch.full((n, 25, 25, 64), -5.0, dtype=torch.float16, device=DEV)
    ops.reid_stem_direct(rd, frames, fh, fw, rois, wpd, out)
    # the round-2 pair of kernels on the same inputs
    wp, bp = ops.pack_reid_stem_weights(w.to(DEV), b.to(DEV))
    patches = torch.zeros(n, 25, 128, 32, dtype=torch.float16, device=DEV)
    out2 = torch.full((n, 25, 25, 64), -5.0, dtype=torch.float16, device=DEV)
    ops.roi_stem_patches(rd, frames, fh, fw, rois, patches)
    ops.reid_stem_pool(patches, wp, bp, out2, n)
    torch.cuda.synchronize()
    assert tuple(lib_fault()) == (0, 0, 0, 0)
    conv = F.conv2d(x[..., :3].float().cpu().permute(0, 3, 1, 2), wp[:, :27].float().cpu().view(64, 3, 3, 3).permute(0, 3, 1, 2), b, 1, 1)
    ref = F.max_pool2d(F.relu(conv).half().float(), 3, 2, 1).permute(0, 2, 3, 1)
    got = out.float().cpu()
    assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    # same operands, same products; only the position of the bias in the fp32 sum differs
    assert (got - out2.float().cpu()).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_reid_direct_stem_train_mode_bn_matches_torch(lib):
    """vcb_reid_stem_direct_stats + vcb_bn_seg_finalize (bias = NULL: it is inside the statistics) + vcb_reid_stem_direct_bn against
    F.batch_norm(training=True) per segment on the same fp16 crop; 300 crops so that CTAs own several crops and segment boundaries
    fall inside a CTA's range."""
    import torch.nn.functional as F
    from vehicle_counting_b200 import ops
    n, nb = 300, 320
    seg_sizes = [5, 1, 120, 64, 110]
    rng, frames, rois_np, rd = _stem_case(61, nb)
    fh, fw = 240, 320
    rois_np[n:] = 0
    rois = torch.from_numpy(rois_np).to(DEV)
    x = torch.zeros(nb, 50, 50, 4, dtype=torch.float16, device=DEV)
    ops.roi_resize_norm(rd, frames, fh, fw, rois, x)
    g = torch.Generator().manual_seed(4)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(64, generator=g) * 0.2
    gamma = torch.rand(64, generator=g) + 0.5
    beta = torch.randn(64, generator=g) * 0.1
    wpd = ops.pack_reid_stem_weights_direct(w.to(DEV), b.to(DEV))
    nseg = len(seg_sizes)
    soc = np.full(nb, nseg, np.int32); soc[:n] = np.repeat(np.arange(nseg), seg_sizes)
    cnt = np.zeros(nseg + 1, np.int32); cnt[:nseg] = seg_sizes; cnt[nseg] = nb - n
    soc_d, cnt_d = torch.from_numpy(soc).to(DEV), torch.from_numpy(cnt).to(DEV)
    sums = torch.zeros(nseg + 1, 64, 2, dtype=torch.float64, device=DEV)
    affine = torch.zeros(nseg + 1, 64, 2, dtype=torch.float32, device=DEV)
    out = torch.full((nb, 25, 25, 64), -5.0, dtype=torch.float16, device=DEV)
    ops.reid_stem_direct_stats(rd, frames, fh, fw, rois, wpd, soc_d, sums)
    ops.bn_seg_finalize(sums, cnt_d, nseg + 1, 64, 2500, gamma.to(DEV), beta.to(DEV), None, 1e-5, affine)
    ops.reid_stem_direct_bn(rd, frames, fh, fw, rois, wpd, affine, soc_d, out)
    torch.cuda.synchronize()
    assert tuple(lib_fault()) == (0, 0, 0, 0)
    wk = wpd[:, :27].float().cpu().view(64, 3, 3, 3).permute(0, 3, 1, 2)
    conv = F.conv2d(x[:n, ..., :3].float().cpu().permute(0, 3, 1, 2), wk, b, 1, 1)
    got_sums = sums.cpu()
    off = 0
    for si, k in enumerate(seg_sizes):
        c_ = conv[off:off + k].double()
        np.testing.assert_allclose(got_sums[si, :, 0].numpy(), c_.sum((0, 2, 3)).numpy(), rtol=2e-4, atol=5e-2)
        np.testing.assert_allclose(got_sums[si, :, 1].numpy(), (c_ * c_).sum((0, 2, 3)).numpy(), rtol=2e-4, atol=5e-2)
        y = F.relu(F.batch_norm(conv[off:off + k], None, None, gamma, beta, True, 0.0, 1e-5))
        ref = F.max_pool2d(y.half().float(), 3, 2, 1).permute(0, 2, 3, 1)
        err = (out[off:off + k].float().cpu() - ref).abs().max().item()
        assert err <= 4e-3 * max(1.0, ref.abs().max().item()), (si, k, err)
        off += k


@pytest.mark.parametrize("c,h,n,seg_sizes,res,pool", [(64, 25, 70, [5, 1, 40, 24], True, False), (512, 4, 300, [64, 64, 64, 64, 44], True, False),
                                                     (256, 7, 130, [130], False, False), (64, 50, 9, [4, 5], False, True)])
def test_bn_seg_stats_finalize_apply_match_torch(lib, c, h, n, seg_sizes, res, pool):
    """train-mode BatchNorm over segments (one reference Extractor call each): vcb_bn_seg_stats_f16 -> vcb_bn_seg_finalize ->
    vcb_bn_seg_apply_f16 against F.batch_norm(training=True) per segment on the same fp16 tensor; blocks that span several crops
    and several segments, padding crops, residual add, the pooled stem form."""
    import torch.nn.functional as F
    from vehicle_counting_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(c + h)
    nb = n + 7                                                 # padding crops -> segment index len(seg_sizes)
    x = (torch.randn(nb, h, h, c, generator=g) * 1.5 + torch.randn(c, generator=g)).half()
    r = torch.randn(nb, h, h, c, generator=g).half() if res else None
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.2
    nseg = len(seg_sizes)
    soc = np.full(nb, nseg, np.int32); soc[:n] = np.repeat(np.arange(nseg), seg_sizes)
    cnt = np.zeros(nseg + 1, np.int32); cnt[:nseg] = seg_sizes; cnt[nseg] = nb - n
    xd, soc_d, cnt_d = x.to(DEV), torch.from_numpy(soc).to(DEV), torch.from_numpy(cnt).to(DEV)
    sums = torch.zeros(nseg + 1, c, 2, dtype=torch.float64, device=DEV)
    aff = torch.zeros(nseg + 1, c, 2, dtype=torch.float32, device=DEV)
    ho = (h + 1) // 2 if pool else h
    y = torch.zeros(nb, ho, ho, c, dtype=torch.float16, device=DEV)
    ops.bn_seg_stats_f16(xd, c, h * h, nb, soc_d, sums)
    ops.bn_seg_finalize(sums, cnt_d, nseg + 1, c, h * h, gamma.to(DEV), beta.to(DEV), None, 1e-5, aff)
    ops.bn_seg_apply_f16(xd, c, h, h, nb, soc_d, aff, None if r is None else r.to(DEV), c if res else 0, L.ACT_RELU, 1 if pool else 0, y, c)
    torch.cuda.synchronize()
    off = 0
    for si, k in enumerate(seg_sizes):
        xs = x[off:off + k].float().permute(0, 3, 1, 2)
        np.testing.assert_allclose(sums[si, :, 0].cpu().numpy(), xs.double().sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-2)
        np.testing.assert_allclose(sums[si, :, 1].cpu().numpy(), (xs.double() ** 2).sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-2)
        ref = F.batch_norm(xs, None, None, gamma, beta, True, 0.0, 1e-5)
        if res:
            ref = ref + r[off:off + k].float().permute(0, 3, 1, 2)
        ref = F.relu(ref)
        if pool:
            ref = F.max_pool2d(ref, 3, 2, 1)
        err = (y[off:off + k].float().cpu() - ref.permute(0, 2, 3, 1)).abs().max().item()
        assert err <= 3e-3 * max(1.0, ref.abs().max().item()), (si, err)
        off += k


def lib_fault():
    from vehicle_counting_b200 import _lib as L
    return L.last_fault()


def test_letterbox_half_is_bit_identical_to_cv2(lib):
    """vcb_letterbox_half_u8 against upstream's host letterbox (cv2.resize INTER_LINEAR + copyMakeBorder 114) for the exact 2x
    ratio the reference hits on 1280x720 video (AutoShape size=640 -> 384x640), and through the YoloBackbone adapter."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import ops
    rng = np.random.default_rng(12)
    for (h0, w0, h1, w1) in ((720, 1280, 384, 640), (96, 160, 64, 80), (64, 64, 32, 32)):
        frames = rng.integers(0, 256, (3, h0, w0, 3), dtype=np.uint8)
        want = np.stack([Y.letterbox(f, (h1, w1)) for f in frames])
        dh, dw = (h1 - h0 // 2) / 2, (w1 - w0 // 2) / 2
        top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
        dst = torch.full((3, h1, w1, 3), 7, dtype=torch.uint8, device=DEV)
        ops.letterbox_half(torch.from_numpy(frames).to(DEV), 3, h0, w0, dst, h1, w1, top, left, 114)
        np.testing.assert_array_equal(dst.cpu().numpy(), want)


def test_adapter_device_letterbox_matches_host_letterbox(lib):
    """YoloBackbone.detect on 1280x720 frames: the device letterbox path and the host (cv2) path give identical rows."""
    from vehicle_counting_b200.networks import yolo as NY
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    rng = np.random.default_rng(13)
    imgs = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(2)]
    net = NY.YoloBackbone(None, 0.45, 0.25, 300, state_dict=synth_yolov5_state_dict("yolov5n", seed=0, obj_bias=1.0))
    det_dev, cnt_dev = net.detect_raw(imgs)
    det_dev, cnt_dev = det_dev.copy(), cnt_dev.copy()
    eng = net._engine(2, 384, 640)
    frames_dev = eng.frames.clone()
    lb = [NY._letterbox(im, (384, 640)) for im in imgs]           # host path: already at the inference shape, no resize left
    eng.set_scale([(720, 1280)] * 2)
    eng.upload(torch.from_numpy(np.stack(lb)).pin_memory()); eng.forward()
    det_host, cnt_host = eng.download()
    assert torch.equal(eng.frames, frames_dev)                    # the network saw the same bytes on both paths
    np.testing.assert_array_equal(cnt_dev, cnt_host)
    np.testing.assert_array_equal(det_dev, det_host)


def test_letterbox_bilinear_is_bit_identical_to_cv2(lib):
    """vcb_letterbox_bilinear_u8 (OpenCV's 11-bit fixed-point bilinear restated on the device) against upstream's host letterbox
    (cv2.resize INTER_LINEAR + copyMakeBorder 114): down-scaling, up-scaling, non-integer ratios, pad-only; and through the
    YoloBackbone adapter (same bytes in the engine's input as with the host path)."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import ops
    from vehicle_counting_b200.networks import yolo as NY
    rng = np.random.default_rng(21)
    for (h0, w0, size) in ((1080, 1920, 640), (720, 1280, 1280), (480, 640, 640), (200, 300, 320), (333, 517, 640), (100, 150, 480)):
        g = size / max(h0, w0)
        h1, w1 = (NY._make_divisible(v * g, 32) for v in (h0, w0))
        frames = rng.integers(0, 256, (2, h0, w0, 3), dtype=np.uint8)
        want = np.stack([Y.letterbox(f, (h1, w1)) for f in frames])
        r = min(h1 / h0, w1 / w0)
        nw, nh = int(round(w0 * r)), int(round(h0 * r))
        top, left = int(round((h1 - nh) / 2 - 0.1)), int(round((w1 - nw) / 2 - 0.1))
        xt = torch.from_numpy(NY.cv2_linear_table(nw, w0, False)).to(DEV)
        yt = torch.from_numpy(NY.cv2_linear_table(nh, h0, True)).to(DEV)
        dst = torch.full((2, h1, w1, 3), 7, dtype=torch.uint8, device=DEV)
        ops.letterbox_bilinear(torch.from_numpy(frames).to(DEV), 2, h0, w0, dst, h1, w1, top, left, nh, nw, xt, yt, 114)
        np.testing.assert_array_equal(dst.cpu().numpy(), want, err_msg=str((h0, w0, h1, w1)))
    # adapter: 1080p frames reach the engine as the host letterbox would deliver them
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    imgs = [rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8) for _ in range(2)]
    net = NY.YoloBackbone(None, 0.45, 0.25, 300, state_dict=synth_yolov5_state_dict("yolov5n", seed=0, obj_bias=1.0))
    net.detect_raw(imgs)
    eng = net._engine(2, 384, 640)
    want = np.stack([NY._letterbox(im, (384, 640)) for im in imgs])
    np.testing.assert_array_equal(eng.frames.cpu().numpy(), want)


@pytest.mark.parametrize("n,h,cin,cout,k,s,seg_sizes", [(37, 25, 64, 64, 3, 1, [5, 1, 20, 11]), (40, 13, 128, 128, 3, 1, [40]),
                                                        (300, 4, 256, 512, 3, 1, [64, 64, 64, 64, 44]), (23, 13, 64, 40, 1, 2, [3, 20]),
                                                        (130, 7, 128, 256, 3, 2, [1] * 130)])
def test_conv_epilogue_bn_statistics_match_the_statistics_kernel(lib, n, h, cin, cout, k, s, seg_sizes):
    """vcb_conv2d_fwd_stats: the convolution's epilogue accumulates the train-mode BatchNorm statistics of the fp16 tensor it stores;
    they must equal what vcb_bn_seg_stats_f16 computes from that tensor (128-row and CTA-pair kernels, tiles that straddle segments,
    one-crop segments, a cout that is not a multiple of 64, the M tail)."""
    from vehicle_counting_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(n + cout)
    x = torch.randn(n, h, h, cin, generator=g).half().to(DEV)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = (torch.randn(cout, generator=g) * 0.3).to(DEV)
    p = k // 2
    cp = (cout + 7) // 8 * 8
    d = ops.make_conv_desc(n, h, h, cin, cout, k, s, p, cout_pitch=cp, act=L.ACT_NONE)
    ho, wo = ops.conv_out_hw(d)
    wp, bp = ops.pack_conv_weights(d, w, b)
    nseg = len(seg_sizes)
    soc = torch.from_numpy(np.repeat(np.arange(nseg), seg_sizes).astype(np.int32)).to(DEV)
    y1 = torch.zeros(n, ho, wo, cp, dtype=torch.float16, device=DEV)
    y2 = torch.zeros_like(y1)
    sums = torch.zeros(nseg + 1, cout, 2, dtype=torch.float64, device=DEV)
    ops.conv2d(d, x, wp, bp, y1)
    ops.conv2d_stats(d, x, wp, bp, y2, soc, sums)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    yf = y1[..., :cout].double().cpu()
    off = 0
    for si, kk in enumerate(seg_sizes):
        blk = yf[off:off + kk]
        np.testing.assert_allclose(sums[si, :, 0].cpu().numpy(), blk.sum((0, 1, 2)).numpy(), rtol=1e-5, atol=5e-3)
        np.testing.assert_allclose(sums[si, :, 1].cpu().numpy(), (blk * blk).sum((0, 1, 2)).numpy(), rtol=1e-5, atol=5e-3)
        off += kk
    assert (sums[nseg] == 0).all()


def test_h2d_frames_inplace_takes_pinned_frames_and_declines_pageable_ones(lib):
    """vcb_h2d_frames_inplace / hostcopy.upload_frames: frames that live in page-locked memory (views of one pinned tensor: merged
    into one copy; separately pinned tensors: one copy each) reach the device without the gather; a pageable frame in the list makes
    the call decline (0) and upload_frames falls back to its pinned staging buffer.  Same bytes on the device either way."""
    from vehicle_counting_b200 import hostcopy
    rng = np.random.default_rng(11)
    n, h, w = 6, 48, 80
    stream = torch.cuda.current_stream()
    block = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).pin_memory()
    views = [block.numpy()[i] for i in range(n)]
    dev = torch.zeros(n, h, w, 3, dtype=torch.uint8, device=DEV)
    assert hostcopy.upload_inplace(dev, views, stream) is True
    torch.cuda.synchronize()
    assert torch.equal(dev.cpu(), block)
    singles_t = [torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).pin_memory() for _ in range(n)]
    singles = [t.numpy() for t in singles_t]
    dev.zero_()
    assert hostcopy.upload_inplace(dev, singles, stream) is True
    torch.cuda.synchronize()
    assert torch.equal(dev.cpu(), torch.stack(singles_t))
    mixed = list(singles)
    mixed[3] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)             # pageable
    dev.zero_()
    assert hostcopy.upload_inplace(dev, mixed, stream) is False
    torch.cuda.synchronize()
    assert int(dev.sum()) == 0                                             # nothing was queued
    staging = torch.empty(n, h, w, 3, dtype=torch.uint8).pin_memory()
    hostcopy.upload_frames(staging, dev, mixed, stream)
    torch.cuda.synchronize()
    assert torch.equal(dev.cpu(), torch.from_numpy(np.stack(mixed)))
    assert hostcopy.upload_inplace(dev, [v[:, :, ::-1] for v in views], stream) is False      # BGR views are not contiguous


@pytest.mark.parametrize("c,h,n,seg_sizes,res", [(64, 25, 70, [5, 1, 40, 24], "plain"), (128, 13, 130, [64, 64, 2], "bn"), (512, 4, 300, [64, 64, 64, 64, 44], "bn"),
                                                 (256, 7, 40, [40], "none")])
def test_bn_fused_apply_equals_finalize_plus_apply(lib, c, h, n, seg_sizes, res):
    """vcb_bn_seg_apply_fused_f16 (scale / shift derived in the kernel; the residual optionally a pre-BN tensor normalised on the fly --
    the downsample branch of a BasicBlock, model.py:33-37) against the two-step form it replaces: bit-identical without a BN'd
    residual, and against F.batch_norm(training=True) per segment with one."""
    import torch.nn.functional as F
    from vehicle_counting_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(c + n)
    nseg = len(seg_sizes)
    x = (torch.randn(n, h, h, c, generator=g) * 1.7 + 0.3).half().to(DEV)
    r = (torch.randn(n, h, h, c, generator=g) * 0.9 - 0.2).half().to(DEV)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.1).to(DEV)
    gr, br = (torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.1).to(DEV)
    soc = torch.from_numpy(np.repeat(np.arange(nseg), seg_sizes).astype(np.int32)).to(DEV)
    cnt = torch.tensor(seg_sizes + [0], dtype=torch.int32, device=DEV)
    sums = torch.zeros(nseg + 1, c, 2, dtype=torch.float64, device=DEV)
    rsums = torch.zeros_like(sums)
    ops.bn_seg_stats_f16(x, c, h * h, n, soc, sums)
    ops.bn_seg_stats_f16(r, c, h * h, n, soc, rsums)
    y = torch.zeros(n, h, h, c, dtype=torch.float16, device=DEV)
    ops.bn_seg_apply_fused_f16(x, c, h * h, n, soc, cnt, sums, gamma, beta, 1e-5, None if res == "none" else r, c, L.ACT_RELU, y, c,
                               res_sums=rsums if res == "bn" else None, res_gamma=gr if res == "bn" else None, res_beta=br if res == "bn" else None)
    torch.cuda.synchronize()
    if res != "bn":
        aff = torch.zeros(nseg + 1, c, 2, dtype=torch.float32, device=DEV)
        y2 = torch.zeros_like(y)
        ops.bn_seg_finalize(sums, cnt, nseg + 1, c, h * h, gamma, beta, None, 1e-5, aff)
        ops.bn_seg_apply_f16(x, c, h, h, n, soc, aff, None if res == "none" else r, c, L.ACT_RELU, 0, y2, c)
        torch.cuda.synchronize()
        assert torch.equal(y, y2)
    off = 0
    for k in seg_sizes:
        xs = x[off:off + k].float().cpu().permute(0, 3, 1, 2)
        want = F.batch_norm(xs, None, None, gamma.cpu(), beta.cpu(), True, 0.0, 1e-5)
        if res == "plain":
            want = want + r[off:off + k].float().cpu().permute(0, 3, 1, 2)
        elif res == "bn":
            want = want + F.batch_norm(r[off:off + k].float().cpu().permute(0, 3, 1, 2), None, None, gr.cpu(), br.cpu(), True, 0.0, 1e-5)
        want = F.relu(want).permute(0, 2, 3, 1)
        err = (y[off:off + k].float().cpu() - want).abs().max().item()
        assert err <= 4e-3 * max(1.0, want.abs().max().item()), (k, err)
        off += k
