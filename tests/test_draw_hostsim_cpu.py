"""The device code of leaffliction_b200/csrc/lfx_draw.cu compiled by g++ as plain C++ (tests/hostsim/lfx_common.cuh stubs the
CUDA qualifiers and intrinsics; a block is run as ONE thread, so every loop over threadIdx.x covers the whole range) and
compared with the golden overlays of the reference and with the oracle.  This checks the kernel's integer / fixed-point
logic on a box without a GPU; the multi-thread execution is what tests/test_gpu_draw.py covers on the B200.
Test infrastructure only: nothing in the package can reach this build."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import spec_draw as sd
from oracle import spec_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "hostsim")

DRIVERS = r'''
extern "C" void host_analyze_overlay(const uint8_t* rgb, const int32_t* points, const int32_t* counts, const int32_t* rec_i,
    const int32_t* hull, const uint8_t* edges, const uint8_t* mask, uint8_t* overlay, int B, int H, int W, int max_pts, int max_hull) {
    for (int b = 0; b < B; ++b) { blockIdx.x = b; k_analyze_overlay(rgb, points, counts, rec_i, hull, edges, mask, overlay, H, W, max_pts, max_hull); }
}
extern "C" void host_draw_rectangles(const uint8_t* rgb, const int32_t* info, uint8_t* vis, int B, int H, int W, uint32_t col, int thickness) {
    for (int b = 0; b < B; ++b) { blockIdx.x = b; k_draw_rectangles(rgb, info, vis, H, W, col, thickness); }
}
extern "C" void host_draw_primitives(uint8_t* img, const int32_t* prims, const int32_t* counts, int B, int H, int W, int max_prims) {
    for (int b = 0; b < B; ++b) { blockIdx.x = b; k_draw_primitives(img, prims, counts, H, W, max_prims); }
}
extern "C" void host_thick_line(uint8_t* img, int H, int W, int x0, int y0, int x1, int y1, uint32_t col, int th, int flags) {
    Img im = {img, H, W}; thick_line(im, x0, y0, x1, y1, col, th, flags); }
extern "C" void host_line_aa(uint8_t* img, int H, int W, int x0, int y0, int x1, int y1, uint32_t col) {
    Img im = {img, H, W}; line_aa_block(im, x0, y0, x1, y1, col); }
extern "C" void host_thick_line_shared(uint8_t* img, int H, int W, int x0, int y0, int x1, int y1, uint32_t col, int th, int flags, int ts) {
    Img im = {img, H, W}; for (int t0 = ts - 1; t0 >= 0; --t0) thick_line(im, x0, y0, x1, y1, col, th, flags, t0, ts); }
extern "C" void host_line_aa_shared(uint8_t* img, int H, int W, int x0, int y0, int x1, int y1, uint32_t col, int ts) {
    Img im = {img, H, W}; blockDim.x = ts;
    for (int t0 = ts - 1; t0 >= 0; --t0) { threadIdx.x = t0; line_aa_block(im, x0, y0, x1, y1, col); }
    blockDim.x = 1; threadIdx.x = 0; }
extern "C" void host_circle(uint8_t* img, int H, int W, int x, int y, int r, uint32_t col) {
    Img im = {img, H, W}; circle_filled(im, x, y, r, col); }
'''


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = open(os.path.join(ROOT, "leaffliction_b200", "csrc", "lfx_draw.cu")).read()
    body = src[:src.index("}  // namespace\n")] + "}  // namespace\n" + DRIVERS
    d = tmp_path_factory.mktemp("drawsim")
    shutil.copy(os.path.join(SIM, "lfx_common.cuh"), d / "lfx_common.cuh")
    (d / "draw_host.cpp").write_text(body)
    so = d / "libdrawhost.so"
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so), str(d / "draw_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(so))


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def u32(c):
    return int(c[0]) | int(c[1]) << 8 | int(c[2]) << 16


def device_layout_record(contour):
    """rec_i32 / hull buffers as lfx_analyze_record lays them out (hull counter-clockwise from its top-most vertex)."""
    pts = np.ascontiguousarray(np.asarray(contour).reshape(-1, 2), np.int32)
    rec = sd.overlay_record(pts.reshape(-1, 1, 2))
    hull = spec_mask.convex_hull_points(pts)
    k = min(range(len(hull)), key=lambda t: (hull[t][1], hull[t][0]))
    hull = np.concatenate([hull[k:], hull[:k]]).astype(np.int32)
    ri = np.zeros((1, 24), np.int32)
    ri[0, 0], ri[0, 1] = 1, len(pts)
    ri[0, 2:4] = rec["centroid"]
    for e, key in enumerate(("left", "right", "top", "bottom")):
        ri[0, 4 + 2 * e:6 + 2 * e] = rec[key]
    ri[0, 12] = len(hull)
    (a0, a1), (b0, b1) = rec["axes"]
    ri[0, 14:22] = [*a0, *a1, *b0, *b1]
    pbuf = np.zeros((1, len(pts) + 5, 2), np.int32)
    pbuf[0, :len(pts)] = pts
    hbuf = np.zeros((1, len(hull) + 3, 2), np.int32)
    hbuf[0, :len(hull)] = hull
    return pbuf, np.array([len(pts)], np.int32), ri, hbuf


def test_primitives_host_sim_vs_oracle(sim):
    rng = np.random.default_rng(9)
    H, W = 40, 56
    for t in range(150):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        lo, hi = (-9, 66) if t % 2 else (0, 40)
        p0 = (int(rng.integers(lo, hi)), int(rng.integers(lo, hi)))
        p1 = p0 if t % 7 == 0 else (int(rng.integers(lo, hi)), int(rng.integers(lo, hi)))
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        for th in (2, 3):
            a = img.copy(); sd.thick_line(a, p0, p1, col, th, 3)
            b = img.copy(); sim.host_thick_line(P(b), H, W, *p0, *p1, u32(col), th, 3)
            assert np.array_equal(a, b), ("thick", th, p0, p1)
            for ts in (3, 32):      # the work of one primitive dealt out to `ts` threads (run one after the other, last first)
                b = img.copy(); sim.host_thick_line_shared(P(b), H, W, *p0, *p1, u32(col), th, 3, ts)
                assert np.array_equal(a, b), ("thick shared", th, ts, p0, p1)
        a = img.copy(); sd.line_aa_px(a, p0, p1, col)
        b = img.copy(); sim.host_line_aa(P(b), H, W, *p0, *p1, u32(col))
        assert np.array_equal(a, b), ("aa", p0, p1)
        b = img.copy(); sim.host_line_aa_shared(P(b), H, W, *p0, *p1, u32(col), 5)
        assert np.array_equal(a, b), ("aa shared", p0, p1)
        a = img.copy(); sd.circle_filled(a, p0, 3, col)
        b = img.copy(); sim.host_circle(P(b), H, W, *p0, 3, u32(col))
        assert np.array_equal(a, b), ("circle", p0)
        x, y = int(rng.integers(0, W - 2)), int(rng.integers(0, H - 2))
        w, h = int(rng.integers(1, W - x + 1)), int(rng.integers(1, H - y + 1))
        a = img.copy(); sd.rectangle2(a, x, y, w, h)
        b = np.zeros_like(img)
        sim.host_draw_rectangles(P(img), P(np.array([[1, x, y, w, h, 0, 0, 0]], np.int32)), P(b), 1, H, W, u32((255, 0, 0)), 2)
        assert np.array_equal(a, b), ("rect", x, y, w, h)


def test_overlays_host_sim_vs_golden(sim):
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_draw_v1.npz"))
    keys = sorted({k.rsplit("_rgb", 1)[0] for k in g.files if k.endswith("_rgb")})
    for k in keys:
        rgb, mask = np.ascontiguousarray(g[k + "_rgb"]), np.ascontiguousarray(g[k + "_mask"])
        edges = np.ascontiguousarray(g[k + "_edges"])
        H, W = rgb.shape[:2]
        pbuf, cnt, ri, hbuf = device_layout_record(g[k + "_contour"])
        out = np.zeros_like(rgb)
        sim.host_analyze_overlay(P(rgb), P(pbuf), P(cnt), P(ri), P(hbuf), P(edges), P(mask), P(out), 1, H, W, pbuf.shape[1], hbuf.shape[1])
        assert np.array_equal(out, g[k + "_analyze"]), k
        vis = np.zeros_like(rgb)
        box = g[k + "_roi_box"]
        sim.host_draw_rectangles(P(rgb), P(np.array([[1, *box, 0, 0, 0]], np.int32)), P(vis), 1, H, W, u32((255, 0, 0)), 2)
        assert np.array_equal(vis, g[k + "_roi_vis"]), k


def random_primitives(rng, B, P, H, W, margin=8):
    """[B,P,8] rows for lfx_draw_primitives: every kind, end points inside and outside the image, overlapping on purpose."""
    prims = np.zeros((B, P, 8), np.int32)
    prims[..., 0] = rng.integers(1, 6, size=(B, P))
    prims[..., 1] = rng.integers(-margin, W + margin, size=(B, P))
    prims[..., 2] = rng.integers(-margin, H + margin, size=(B, P))
    prims[..., 3] = rng.integers(-margin, W + margin, size=(B, P))
    prims[..., 4] = rng.integers(-margin, H + margin, size=(B, P))
    prims[..., 5] = rng.integers(0, 1 << 24, size=(B, P))
    kind = prims[..., 0]
    prims[..., 6] = np.where(kind == 3, rng.integers(0, 7, size=(B, P)), rng.integers(2, 5, size=(B, P)))
    mk = kind == 5
    prims[..., 3] = np.where(mk, rng.integers(4, 21, size=(B, P)), prims[..., 3])     # markerSize
    return prims


def test_primitive_lists_host_sim_vs_oracle(sim):
    rng = np.random.default_rng(21)
    B, P, H, W = 12, 9, 40, 56
    imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
    prims = random_primitives(rng, B, P, H, W)
    counts = rng.integers(0, P + 1, size=B).astype(np.int32)
    got = imgs.copy()
    sim.host_draw_primitives(got.ctypes.data_as(C.c_void_p), prims.ctypes.data_as(C.c_void_p),
                             counts.ctypes.data_as(C.c_void_p), B, H, W, P)
    for b in range(B):
        exp = sd.draw_primitives(imgs[b].copy(), prims[b, :counts[b]])
        assert np.array_equal(got[b], exp), (b, prims[b, :counts[b]].tolist())
