"""The device code of leaffliction_b200/csrc/lfx_contour.cu (contour trace: global-memory walk and the shared-memory bit-plane
variant) compiled by g++ as plain C++ (tests/hostsim/lfx_common.cuh; a block runs as ONE thread) against the oracle's border
following and cv2.findContours.  Checks the walk, the bit packing and the direction encoding without a GPU; the GPU suite covers
the real launch.  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from leaffliction_b200 import synth
from oracle import spec_contour

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVERS = r'''
extern "C" void host_record(const int32_t* points, const int32_t* counts, int32_t* rec_i, double* rec_f, int32_t* hull, int32_t* ws,
                            int B, int H, int max_pts, int max_hull, int smem) {
    for (int b = 0; b < B; ++b) {
        blockIdx.x = b; threadIdx.x = 0; blockDim.x = 1;
        if (smem) k_analyze_record<true>(points, counts, nullptr, rec_i, rec_f, hull, nullptr, B, H, max_pts, max_hull);
        else k_analyze_record<false>(points, counts, nullptr, rec_i, rec_f, hull, ws, B, H, max_pts, max_hull);
    }
}
extern "C" void host_trace(const uint8_t* mask, const int32_t* info, int32_t* points, int32_t* counts, long long* sums,
                           int B, int H, int W, int max_pts, int bits) {
    for (int b = 0; b < B; ++b) {
        if (bits) { blockIdx.x = b; threadIdx.x = 0; blockDim.x = 1; k_trace_contour_bits(mask, info, points, counts, sums, H, W, max_pts); }
        else { blockIdx.x = b; threadIdx.x = 0; blockDim.x = 1; k_trace_contour(mask, info, points, counts, sums, B, H, W, max_pts); }
    }
}
'''


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = open(os.path.join(ROOT, "leaffliction_b200", "csrc", "lfx_contour.cu")).read()
    body = src[:src.index("}  // namespace\n")] + "}  // namespace\n" + DRIVERS
    body = body.replace("extern __shared__ uint32_t s_bits[];", "static uint32_t s_bits[1 << 16];")
    body = body.replace("extern __shared__ int32_t s_ws[];", "static int32_t s_ws[1 << 16];")
    d = tmp_path_factory.mktemp("contoursim")
    shutil.copy(os.path.join(ROOT, "tests", "hostsim", "lfx_common.cuh"), d / "lfx_common.cuh")
    (d / "contour_host.cpp").write_text(body)
    so = d / "libcontourhost.so"
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so), str(d / "contour_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(so))


def _masks():
    rng = np.random.default_rng(7)
    out = []
    for (H, W) in ((64, 64), (40, 70), (96, 128)):
        for k in range(6):
            m = np.zeros((H, W), np.uint8)
            yy, xx = np.mgrid[0:H, 0:W]
            for _ in range(int(rng.integers(1, 5))):
                cy, cx = rng.integers(5, H - 5), rng.integers(5, W - 5)
                ry, rx = rng.integers(2, H // 2), rng.integers(2, W // 2)
                m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 255
            if k % 2:
                m[rng.random((H, W)) < 0.1] = 0
            if k == 3:
                m[:] = 255                       # the whole image: the border runs along the image edge
            if k == 4:
                m[:] = 0
                m[H // 2, W // 3] = 255          # an isolated pixel
            out.append(m)
    return out


def _first_component(m):
    """(info row, mask of the 8-connected component of the first raster pixel) -- what lfx_make_mask hands to the trace."""
    from oracle import spec_filters as sf
    ys, xs = np.nonzero(m)
    if len(ys) == 0:
        return np.zeros(8, np.int32), m
    lab = sf.label8(m)
    comp = (lab == lab[ys[0], xs[0]]).astype(np.uint8) * 255
    yy, xx = np.nonzero(comp)
    info = np.zeros(8, np.int32)
    info[0], info[1], info[2] = 1, xx.min(), yy.min()
    info[3], info[4] = xx.max() - xx.min() + 1, yy.max() - yy.min() + 1
    info[7] = int(xs[0]) << 8
    return info, comp


@pytest.mark.parametrize("bits", [0, 1])
def test_trace_host_sim_vs_oracle_and_cv2(sim, bits):
    cv2 = pytest.importorskip("cv2")
    for m in _masks():
        info, comp = _first_component(m)
        H, W = comp.shape
        max_pts = 2048
        pts = np.zeros((1, max_pts, 2), np.int32)
        cnt = np.zeros(1, np.int32)
        sums = np.zeros((1, 3), np.int64)
        comp = np.ascontiguousarray(comp)
        sim.host_trace(comp.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p), pts.ctypes.data_as(C.c_void_p),
                       cnt.ctypes.data_as(C.c_void_p), sums.ctypes.data_as(C.c_void_p), 1, H, W, max_pts, bits)
        if not info[0]:
            assert cnt[0] == 0
            continue
        exp = spec_contour.trace_external(comp, (int(info[7]) >> 8, int(info[2])))
        assert cnt[0] == len(exp) and np.array_equal(pts[0, :cnt[0]], exp[:, 0, :])
        cs, _ = cv2.findContours(comp, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        assert len(cs) == 1 and np.array_equal(cs[0][:, 0, :], pts[0, :cnt[0]])
        M = cv2.moments(cs[0])
        assert abs(sums[0, 0]) / 2 == M["m00"]


def test_record_host_sim_both_variants(sim):
    """k_analyze_record with its scratch in the global workspace and in shared memory: identical records (bit for bit, the
    arithmetic and its order are the same), centroid / extreme points / hull as the oracle has them."""
    cv2 = pytest.importorskip("cv2")
    from oracle import spec_mask
    seen = 0
    for m in _masks():
        info, comp = _first_component(m)
        if not info[0]:
            continue
        H, W = comp.shape
        cs, _ = cv2.findContours(np.ascontiguousarray(comp), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        c = cs[0].astype(np.int32)
        n, max_pts, max_hull = len(c), len(c) + 3, 256
        pts = np.zeros((1, max_pts, 2), np.int32)
        pts[0, :n] = c[:, 0, :]
        cnt = np.array([n], np.int32)
        outs = []
        for smem in (0, 1):
            ri, rf = np.full((1, 24), -7, np.int32), np.full((1, 12), -7.0, np.float64)
            hull = np.zeros((1, max_hull, 2), np.int32)
            ws = np.zeros(6 * H + 8, np.int32)
            sim.host_record(pts.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p),
                            rf.ctypes.data_as(C.c_void_p), hull.ctypes.data_as(C.c_void_p), ws.ctypes.data_as(C.c_void_p),
                            1, H, max_pts, max_hull, smem)
            outs.append((ri.copy(), rf.copy(), hull.copy()))
        for a, b in zip(outs[0], outs[1]):
            assert a.tobytes() == b.tobytes()
        ri, rf, hull = outs[1]
        exp = spec_contour.analyze_record(c)
        assert (ri[0, 2], ri[0, 3]) == tuple(exp["centroid"])
        for e, key in enumerate(("left", "right", "top", "bottom")):
            assert (ri[0, 4 + 2 * e], ri[0, 5 + 2 * e]) == tuple(int(v) for v in exp[key])
        nh = ri[0, 12]
        assert {tuple(p) for p in hull[0, :nh]} == {tuple(int(v) for v in p) for p in spec_mask.convex_hull_points(c[:, 0, :])}
        seen += 1
    assert seen >= 10
