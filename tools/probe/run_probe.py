#!/usr/bin/env python
"""Drive tools/probe/libhpvg_probe.so on a B200: checks every tcgen05 / TMA layout assumption the convolution
kernels make, one hypothesis per test, and prints PASS/FAIL with the max abs error.  Development tool."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libhpvg_probe.so"))


class ProbeOp(ctypes.Structure):
    _fields_ = [("a_desc", ctypes.c_uint64), ("b_desc", ctypes.c_uint64), ("idesc", ctypes.c_uint32),
                ("accumulate", ctypes.c_uint32), ("d_col", ctypes.c_uint32), ("pad", ctypes.c_uint32)]


lib.hpvg_probe_umma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ProbeOp), ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p]
lib.hpvg_probe_umma_time.restype = ctypes.c_longlong
lib.hpvg_probe_umma_time.argtypes = [ctypes.POINTER(ProbeOp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int]
lib.hpvg_probe_tma.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                               ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                               ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                               ctypes.c_void_p, ctypes.c_int]


def to_bf16_bits(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def bf16_round(x):
    return (to_bf16_bits(x).astype(np.uint32) << 16).view(np.float32)


def smem_desc(addr, lbo, sbo, layout=0, base_offset=0):
    d = (addr >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= (layout & 7) << 61
    return d


def idesc_bf16(M, N, a_mn=0, b_mn=0):
    d = (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)
    return d


def run_umma(image_u16, ops, n_cols):
    img = np.ascontiguousarray(image_u16)
    arr = (ProbeOp * len(ops))()
    for i, (a, b, idesc, acc, col) in enumerate(ops):
        arr[i] = ProbeOp(a, b, idesc, acc, col, 0)
    out = np.zeros((128, n_cols), dtype=np.float32)
    rc = lib.hpvg_probe_umma(img.ctypes.data, img.nbytes, arr, len(ops), n_cols, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("probe_umma rc=%d" % rc)
    return out


RESULTS = []


def report(name, got, want, tol=1e-2):
    err = float(np.max(np.abs(got - want)))
    ok = err <= tol * max(1.0, float(np.max(np.abs(want))))
    RESULTS.append((name, ok, err))
    print("%-58s %s  max|err|=%.4g" % (name, "PASS" if ok else "FAIL", err), flush=True)
    return ok


rng = np.random.default_rng(0)


def rnd(*shape):
    return bf16_round(rng.standard_normal(shape).astype(np.float32))


def kmajor_nosw_image(mat):
    """mat [rows][K] -> [K/8][rows][8] bf16 bits (core matrices 8 rows x 16B contiguous)."""
    rows, K = mat.shape
    return to_bf16_bits(mat.reshape(rows, K // 8, 8).transpose(1, 0, 2))


# ------------------------------------------------------------------------------------------ test 1: basic
def test_basic(N=64):
    A = rnd(128, 16)
    B = rnd(N, 16)
    ia = kmajor_nosw_image(A).reshape(-1)      # LBO = 128*16 = 2048, SBO = 128
    ib = kmajor_nosw_image(B).reshape(-1)      # LBO = N*16, SBO = 128
    img = np.concatenate([ia, ib])
    a = smem_desc(0, 2048, 128)
    b = smem_desc(ia.nbytes, N * 16, 128)
    out = run_umma(img, [(a, b, idesc_bf16(128, N), 0, 0)], N)
    return report("T1 no-swizzle K-major M128 N%d K16" % N, out, A @ B.T)


# ------------------------------------------------------------------------------------------ test 2: row-shifted
def test_shift():
    ok = True
    Afull = rnd(136, 16)
    B = rnd(64, 16)
    ia = kmajor_nosw_image(Afull).reshape(-1)  # [2][136][8]; LBO = 136*16
    ib = kmajor_nosw_image(B).reshape(-1)
    img = np.concatenate([ia, ib])
    for dw in (1, 2, 3, 7):
        a = smem_desc(dw * 16, 136 * 16, 128)
        b = smem_desc(ia.nbytes, 64 * 16, 128)
        out = run_umma(img, [(a, b, idesc_bf16(128, 64), 0, 0)], 64)
        ok &= report("T2 no-swizzle start shifted by %d rows (16B each)" % dw, out, Afull[dw:dw + 128] @ B.T)
    return ok


# ------------------------------------------------------------------------------------------ test 3: LBO overlap
def test_lbo_overlap():
    X = rnd(130, 8)  # one 16B chunk per row
    B = rnd(64, 16)
    ia = to_bf16_bits(X).reshape(-1)
    pad = (-ia.nbytes) % 128
    ia = np.concatenate([ia, np.zeros(pad // 2, np.uint16)])
    ib = kmajor_nosw_image(B).reshape(-1)
    img = np.concatenate([ia, ib])
    ok = True
    for lbo_rows in (1, 2):
        a = smem_desc(0, lbo_rows * 16, 128)
        b = smem_desc(ia.nbytes, 64 * 16, 128)
        out = run_umma(img, [(a, b, idesc_bf16(128, 64), 0, 0)], 64)
        want = X[0:128] @ B[:, :8].T + X[lbo_rows:lbo_rows + 128] @ B[:, 8:].T
        ok &= report("T3 LBO = %d*16B (K-groups overlap, tap pairing)" % lbo_rows, out, want)
    return ok


# ------------------------------------------------------------------------------------------ test 4: conv plane, no swizzle
def test_conv_plane_nosw(CG=8):
    C = CG * 8
    X = rnd(18, 10, C)                       # [h][w][c] haloed plane
    Wt = rnd(9, 64, C) * 0.25                # [tap][cout][cin]
    # image [h][cg][w][8]
    ia = to_bf16_bits(X.reshape(18, 10, CG, 8).transpose(0, 2, 1, 3)).reshape(-1)
    KGS, RS = 10 * 16, CG * 10 * 16
    # weights per tap: K-major no-swizzle [cg][cout][8]
    ib = to_bf16_bits(Wt.reshape(9, 64, CG, 8).transpose(0, 2, 1, 3)).reshape(-1)
    boff = (ia.nbytes + 127) // 128 * 128
    img = np.concatenate([ia, np.zeros((boff - ia.nbytes) // 2, np.uint16), ib])
    ops = []
    first = 0
    for dh in range(3):
        for dw in range(3):
            tap = dh * 3 + dw
            if CG >= 2:
                for kk in range(CG // 2):
                    a = smem_desc(dh * RS + (2 * kk) * KGS + dw * 16, KGS, RS)
                    b = smem_desc(boff + tap * 64 * C * 2 + (2 * kk) * 64 * 16, 64 * 16, 128)
                    ops.append((a, b, idesc_bf16(128, 64), 0 if first == 0 else 1, 0))
                    first = 1
    out = run_umma(img, ops, 64)
    want = np.zeros((16, 8, 64), np.float32)
    for dh in range(3):
        for dw in range(3):
            want += np.einsum("hwc,oc->hwo", X[dh:dh + 16, dw:dw + 8], Wt[dh * 3 + dw])
    return report("T4 conv plane 8w x 16h, 9 taps, no-swizzle [h][cg][w][8], CG=%d" % CG, out,
                  want.reshape(128, 64), tol=2e-2)


# ------------------------------------------------------------------------------------------ test 5: SW128
def sw128_image(rows_by_64):
    """rows [R][64] bf16 -> natural 128B rows with 16B chunk j stored at j ^ (r & 7) (base 1024-aligned)."""
    R = rows_by_64.shape[0]
    bits = to_bf16_bits(rows_by_64).reshape(R, 8, 8)
    out = np.zeros_like(bits)
    for r in range(R):
        for j in range(8):
            out[r, j ^ (r & 7)] = bits[r, j]
    return out.reshape(-1)


def test_sw128():
    ok = True
    Afull = rnd(144, 64)
    B = rnd(64, 64)
    ia = sw128_image(Afull)                   # 144 rows * 128 B = 18432 (multiple of 1024)
    ib = sw128_image(B)                       # 64 rows
    img = np.concatenate([ia, ib])
    boff = ia.nbytes
    for shift, bo, tag in ((0, 0, "aligned"), (1, 0, "row+1 base_offset=0"), (1, 1, "row+1 base_offset=1"),
                           (3, 0, "row+3 base_offset=0"), (3, 3, "row+3 base_offset=3"),
                           (8, 0, "row+8 (aligned atom)")):
        ops = []
        for kk in range(4):
            a = smem_desc(shift * 128 + kk * 32, 16, 1024, layout=2, base_offset=bo)
            b = smem_desc(boff + kk * 32, 16, 1024, layout=2)
            ops.append((a, b, idesc_bf16(128, 64), 0 if kk == 0 else 1, 0))
        out = run_umma(img, ops, 64)
        ok &= report("T5 SW128 K-major K64, A %s" % tag, out, Afull[shift:shift + 128] @ B.T, tol=2e-2)
    return ok


def test_conv_plane_sw128():
    X = rnd(18, 10, 64)
    Wt = rnd(9, 64, 64) * 0.25
    ia = sw128_image(X.reshape(180, 64))      # rows are linear (h*10+w); 180*128 = 23040 B
    boff = (ia.nbytes + 1023) // 1024 * 1024
    ib = np.concatenate([sw128_image(Wt[t]) for t in range(9)])
    img = np.concatenate([ia, np.zeros((boff - ia.nbytes) // 2, np.uint16), ib])
    ok = True
    for use_bo in (0, 1):
        ops = []
        first = 0
        for dh in range(3):
            for dw in range(3):
                tap = dh * 3 + dw
                row0 = dh * 10 + dw
                for kk in range(4):
                    a = smem_desc(row0 * 128 + kk * 32, 16, 1280, layout=2, base_offset=(row0 & 7) if use_bo else 0)
                    b = smem_desc(boff + tap * 8192 + kk * 32, 16, 1024, layout=2)
                    ops.append((a, b, idesc_bf16(128, 64), first, 0))
                    first = 1
        out = run_umma(img, ops, 64)
        want = np.zeros((16, 8, 64), np.float32)
        for dh in range(3):
            for dw in range(3):
                want += np.einsum("hwc,oc->hwo", X[dh:dh + 16, dw:dw + 8], Wt[dh * 3 + dw])
        ok &= report("T6 conv plane SW128 [h][w][128B], SBO=1280, base_offset %s" % ("=(row&7)" if use_bo else "=0"),
                     out, want.reshape(128, 64), tol=2e-2)
    return ok


# ------------------------------------------------------------------------------------------ test 7: MN-major
def test_mn_major():
    A = rnd(16, 128)   # [k][m]
    B = rnd(16, 64)    # [k][n]
    # image [mg][k][8]: core matrix = 8 k-rows x 16B (8 consecutive m) ; SBO (MN dir) = 16*16, LBO (K dir) = 128
    ia = to_bf16_bits(A.reshape(16, 16, 8).transpose(1, 0, 2)).reshape(-1)
    ib = to_bf16_bits(B.reshape(16, 8, 8).transpose(1, 0, 2)).reshape(-1)
    img = np.concatenate([ia, ib])
    ok = True
    for (lbo, sbo, tag) in ((128, 256, "LBO=K-dir SBO=MN-dir"), (256, 128, "LBO=MN-dir SBO=K-dir")):
        a = smem_desc(0, lbo, sbo)
        b = smem_desc(ia.nbytes, lbo, sbo)
        out = run_umma(img, [(a, b, idesc_bf16(128, 64, 1, 1), 0, 0)], 64)
        ok_i = report("T7 MN-major A,B no-swizzle (%s)" % tag, out, A.T @ B)
        ok = ok or ok_i
    return ok


def test_m64():
    A = rnd(64, 16)
    B = rnd(64, 16)
    ia = kmajor_nosw_image(A).reshape(-1)
    ib = kmajor_nosw_image(B).reshape(-1)
    img = np.concatenate([ia, ib])
    a = smem_desc(0, 64 * 16, 128)
    b = smem_desc(ia.nbytes, 64 * 16, 128)
    out = run_umma(img, [(a, b, idesc_bf16(64, 64), 0, 0)], 64)
    want = A @ B.T
    # find where rows landed
    lanes = []
    for i in range(64):
        d = np.abs(out - want[i][None, :]).max(axis=1)
        lanes.append(int(np.argmin(d)))
    print("T8 M=64: row->lane map (first 20):", lanes[:20], " rows 16..19 ->", lanes[16:20], "32.. ->", lanes[32:36],
          flush=True)
    return True



# ------------------------------------------------------------------------------------------ test 11: MN-major SW128 (wgrad operands)
def test_mn_sw128():
    """A = x^T, B = gy^T taken straight from channels-last [voxel][64ch] SW128 rows (what TMA writes):
    D[m=ci (+64 for the second x row)][n=co] = sum_k x[row0+k(+P)][ci] * gy[g0+k][co], K = 16 voxels, shifted starts."""
    ok = True
    P = 66                                   # x row pitch in voxels (64 + 2 halo)
    X = rnd(3 * P, 64)                       # 3 image rows of 66 voxels
    GY = rnd(64, 64)                         # one gy row segment of 64 voxels
    ia = sw128_image(X)                      # 198 rows * 128 = 25344 B
    boff = (ia.nbytes + 1023) // 1024 * 1024
    ib = sw128_image(GY)
    img = np.concatenate([ia, np.zeros((boff - ia.nbytes) // 2, np.uint16), ib])
    for (row0, g0) in ((0, 0), (1, 0), (19, 16), (2 + 32, 32), (5, 48)):
        # M=128: block 0 = x rows starting row0, block 1 = the same voxels one image row below (LBO = P*128)
        a = smem_desc(row0 * 128, P * 128, 1024, layout=2)
        b = smem_desc(boff + g0 * 128, 16, 1024, layout=2)
        out = run_umma(img, [(a, b, idesc_bf16(128, 64, 1, 1), 0, 0)], 64)
        want = np.concatenate([X[row0:row0 + 16].T @ GY[g0:g0 + 16], X[row0 + P:row0 + P + 16].T @ GY[g0:g0 + 16]])
        ok &= report("T11 MN-major SW128 M128 (2 x-rows via LBO) row0=%d g0=%d" % (row0, g0), out, want, tol=2e-2)
    # M=64 variant
    a = smem_desc(3 * 128, P * 128, 1024, layout=2)
    b = smem_desc(boff, 16, 1024, layout=2)
    out = run_umma(img, [(a, b, idesc_bf16(64, 64, 1, 1), 0, 0)], 64)
    want = X[3:19].T @ GY[0:16]
    lanes = [(i // 16) * 32 + i % 16 for i in range(64)]
    ok &= report("T11 MN-major SW128 M64 (lane map (i/16)*32+i%16)", out[lanes], want, tol=2e-2)
    return ok

# ------------------------------------------------------------------------------------------ TMA
def run_tma(src_u16, dims, strides_bytes, box, swizzle, coords):
    rank = len(dims)
    d = (ctypes.c_longlong * 5)(*list(dims) + [1] * (5 - rank))
    s = (ctypes.c_longlong * 4)(*list(strides_bytes) + [0] * (4 - len(strides_bytes)))
    b = (ctypes.c_int * 5)(*list(box) + [1] * (5 - rank))
    c = (ctypes.c_int * 5)(*list(coords) + [0] * (5 - rank))
    nbytes = int(np.prod(box)) * 2
    out = np.zeros(nbytes // 2, np.uint16)
    src = np.ascontiguousarray(src_u16)
    rc = lib.hpvg_probe_tma(src.ctypes.data, src.nbytes, 2, rank, d, s, b, swizzle, c, out.ctypes.data, nbytes)
    if rc != 0:
        raise RuntimeError("probe_tma rc=%d" % rc)
    return out


def test_tma_nosw():
    T, H, W, C = 3, 20, 12, 64
    X = rnd(T, H, W, C)
    bits = to_bf16_bits(X)
    dims = (8, W, 8, H, T)
    strides = (C * 2, 16, W * C * 2, H * W * C * 2)
    box = (8, 10, 8, 18, 2)
    ok = True
    for coords in ((0, 1, 0, 1, 0), (0, -1, 0, -1, -1), (0, 5, 0, 7, 2)):
        got = run_tma(bits, dims, strides, box, 0, coords).reshape(2, 18, 8, 10, 8)
        want = np.zeros((2, 18, 8, 10, 8), np.float32)
        for t in range(2):
            for h in range(18):
                for w in range(10):
                    tt, hh, ww = coords[4] + t, coords[3] + h, coords[1] + w
                    if 0 <= tt < T and 0 <= hh < H and 0 <= ww < W:
                        want[t, h, :, w, :] = X[tt, hh, ww].reshape(8, 8)
        gotf = (got.astype(np.uint32) << 16).view(np.float32)
        ok &= report("T9 TMA 5D box->[t][h][cg][w][8] coords=%s" % (coords,), gotf, want, tol=0)
    return ok


def test_tma_sw128():
    T, H, W, C = 3, 20, 12, 64
    X = rnd(T, H, W, C)
    bits = to_bf16_bits(X)
    dims = (C, W, H, T)
    strides = (C * 2, W * C * 2, H * W * C * 2)
    box = (64, 10, 18, 2)
    ok = True
    for coords in ((0, 1, 1, 0), (0, -1, -1, -1)):
        got = run_tma(bits, dims, strides, box, 3, coords).reshape(360, 8, 8)
        want = np.zeros((360, 8, 8), np.float32)
        for t in range(2):
            for h in range(18):
                for w in range(10):
                    tt, hh, ww = coords[3] + t, coords[2] + h, coords[1] + w
                    r = (t * 18 + h) * 10 + w
                    if 0 <= tt < T and 0 <= hh < H and 0 <= ww < W:
                        v = X[tt, hh, ww].reshape(8, 8)
                        for j in range(8):
                            want[r, j ^ (r & 7)] = v[j]
        gotf = (got.astype(np.uint32) << 16).view(np.float32)
        ok &= report("T10 TMA 4D SW128 box->[t][h][w][128B swz] coords=%s" % (coords,), gotf, want, tol=0)
    return ok


# ------------------------------------------------------------------------------------------ timing
def test_timing():
    def t(ops, tag, smem_used=65536, grid=1):
        arr = (ProbeOp * len(ops))()
        for i, (a, b, idesc, acc, col) in enumerate(ops):
            arr[i] = ProbeOp(a, b, idesc, acc, col, 0)
        r = lib.hpvg_probe_umma_time(arr, len(ops), 128, 64, smem_used, grid)
        print("TIME %-64s grid=%3d : %8.2f cycles/MMA" % (tag, grid, r / 1000.0), flush=True)

    for grid in (1, 148):
        # no-swizzle, distinct A per op (conv-like: 4 k-steps x 9 taps)
        KGS, RS = 160, 1280
        ops = []
        for dh in range(3):
            for dw in range(3):
                for kk in range(4):
                    ops.append((smem_desc(dh * RS + 2 * kk * KGS + dw * 16, KGS, RS),
                                smem_desc(32768 + (dh * 3 + dw) * 0 + 2 * kk * 1024, 1024, 128), idesc_bf16(128, 64), 1, 0))
        t(ops, "no-swizzle conv-like M128 N64 K16", grid=grid)
        ops = []
        for dh in range(3):
            for dw in range(3):
                for kk in range(4):
                    ops.append((smem_desc((dh * 10 + dw) * 128 + kk * 32, 16, 1280, 2),
                                smem_desc(32768 + kk * 32, 16, 1024, 2), idesc_bf16(128, 64), 1, 0))
        t(ops, "SW128 conv-like M128 N64 K16", grid=grid)
        ops = [(smem_desc(0, 2048, 128), smem_desc(8192, 2048, 128), idesc_bf16(128, 128), 1, 0)] * 16
        t(ops, "no-swizzle M128 N128 K16", grid=grid)
        ops = [(smem_desc(0, 2048, 128), smem_desc(8192, 2048, 128), idesc_bf16(128, 256), 1, 0)] * 16
        t(ops, "no-swizzle M128 N256 K16 (alloc 128 cols: N wraps, timing only)", grid=grid) if False else None
        ops = [(smem_desc(0, 2048, 128), smem_desc(8192, 256, 128), idesc_bf16(128, 16), 1, 0)] * 16
        t(ops, "no-swizzle M128 N16 K16", grid=grid)
        ops = [(smem_desc(0, 2048, 128), smem_desc(8192, 512, 128), idesc_bf16(128, 32), 1, 0)] * 16
        t(ops, "no-swizzle M128 N32 K16", grid=grid)


TESTS = {
    "basic64": test_basic, "basic16": lambda: test_basic(16), "shift": test_shift, "lbo_overlap": test_lbo_overlap,
    "conv_nosw": test_conv_plane_nosw, "sw128": test_sw128, "conv_sw128": test_conv_plane_sw128,
    "mn_major": test_mn_major, "mn_sw128": test_mn_sw128, "m64": test_m64, "tma_nosw": test_tma_nosw, "tma_sw128": test_tma_sw128,
    "timing": test_timing,
}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        # one test per process: a faulting kernel poisons the CUDA context
        TESTS[sys.argv[1]]()
        bad = [r for r in RESULTS if not r[1]]
        sys.exit(1 if bad else 0)
    import subprocess
    summary = []
    for name in TESTS:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=180)
            sys.stdout.write(p.stdout)
            if p.returncode != 0:
                sys.stdout.write("  [%s rc=%d] %s\n" % (name, p.returncode, p.stderr.strip()[-600:]))
            summary.append((name, p.returncode))
        except subprocess.TimeoutExpired:
            print("  [%s TIMEOUT]" % name)
            summary.append((name, "timeout"))
        sys.stdout.flush()
    print("SUMMARY:", summary)
    sys.exit(0)
