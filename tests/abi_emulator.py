"""TEST INFRASTRUCTURE: a CPU emulation of the C-ABI entry points the captured plans call (include/imfnet_b200.h), so that the
product's HOST orchestration (imfnet_b200/engine.py, batched.py, model/Img_Encoder.py: buffer wiring, column windows of the
concatenations, BatchNorm folding, weight scaling, per-item slots of a batch) can be executed and compared with the oracle in the
`-m "not gpu"` suite.  It says nothing about the CUDA kernels themselves -- those are checked against the oracle on the GPU.

Conventions of the emulation (internal to this file):
  * an "h2 matrix" of C channels (fp16 hi/lo pairs, 4 bytes per channel) is emulated as fp32 in the same bytes: element (row, c) at
    float index row * ld/2 + c.  Column windows (pointer + channel offset * 4 bytes) therefore work as in the product;
  * opaque objects (hash tables, packed weights, projected K/V) live in dictionaries keyed by their buffer address;
  * device-side counts (n_dev) are read at call time, like the kernels do.
Every function follows the semantics documented in include/imfnet_b200.h for the entry point of the same name."""
import ctypes

import numpy as np
from numpy.lib.stride_tricks import as_strided

from imfnet_b200 import _lib

HOST_ONLY = {"imf_conv_first_tc_columns", "imf_conv_first_tc_grid", "imf_last_error", "imf_version", "imf_launch_count", "imf_hash_capacity", "imf_hash_bytes", "imf_device_sm_count"}


def vec(ptr, n, dtype=np.float32):
    n = max(int(n), 0)
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype)


def mat(ptr, rows, cols, ld, dtype=np.float32):
    """rows x cols view with row stride ld (elements); only (rows-1)*ld + cols elements are ever addressed"""
    rows, cols, ld = int(rows), int(cols), int(ld)
    if rows <= 0:
        return np.zeros((0, cols), dtype=dtype)
    base = vec(ptr, (rows - 1) * ld + cols, dtype)
    it = np.dtype(dtype).itemsize
    return as_strided(base, (rows, cols), (ld * it, it))


def count(n_dev, n_max):
    if not n_dev:
        return int(n_max)
    return int(min(int(vec(n_dev, 1, np.int32)[0]), n_max))


def offsets(K):
    r = np.arange(K) - K // 2
    kz, ky, kx = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([kx.ravel(), ky.ravel(), kz.ravel()], axis=1).astype(np.int64)


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    return ((x - mu) / np.sqrt(var + eps) * w + b).astype(np.float32)


def gelu(x):
    from math import sqrt
    import torch
    return (x * 0.5 * (1.0 + torch.erf(torch.from_numpy(x / sqrt(2.0))).numpy())).astype(np.float32)


class Emulator:
    def __init__(self, real):
        self.real = real
        self.calls = []
        self.tables, self.packed, self.kv = {}, {}, {}

    def __getattr__(self, name):
        if name not in _lib.SIGNATURES:
            raise AttributeError(name)
        if name in HOST_ONLY or name.endswith("_bytes"):
            return getattr(self.real, name)
        fn = self.__class__.__dict__.get("do_" + name)
        if fn is None:
            raise NotImplementedError(f"abi_emulator: {name} is not emulated")
        nargs = len(_lib.SIGNATURES[name][1])

        def call(*args):
            assert len(args) == nargs, f"{name}: {len(args)} arguments, the ABI takes {nargs}"
            self.calls.append(name)
            fn(self, *args[:-1])          # the last argument of every launching entry point is the stream
            return 0

        return call

    # ---- coordinates -------------------------------------------------------------------------------------------------
    def do_imf_hash_build(self, coords, n_dev, n_max, table, cap, status):
        n = count(n_dev, n_max)
        C = mat(coords, n, 4, 4, np.int32)
        d = {}
        for i, row in enumerate(map(tuple, C.tolist())):
            assert row not in d, "duplicate coordinate"
            d[row] = i
        self.tables[table] = d

    def do_imf_stride_map(self, coords_in, n_in_dev, n_in_max, stride, table_out, cap, coords_out, n_out_dev, first_idx, ws, ws_bytes,
                          status):
        n = count(n_in_dev, n_in_max)
        C = mat(coords_in, n, 4, 4, np.int32).copy()
        C[:, 1:] = np.floor_divide(C[:, 1:], stride) * stride
        d, rows, first = {}, [], []
        for i, row in enumerate(map(tuple, C.tolist())):
            if row not in d:
                d[row] = len(rows)
                rows.append(row)
                first.append(i)
        out = mat(coords_out, len(rows), 4, 4, np.int32)
        if rows:
            out[:] = np.asarray(rows, dtype=np.int32)
        vec(n_out_dev, 1, np.int32)[0] = len(rows)
        self.tables[table_out] = d
        if first_idx:          # source row of every kept row (imf_stride_map with stride 1 = ME.utils.sparse_quantize(return_index=True))
            vec(first_idx, len(first), np.int32)[:] = np.asarray(first, dtype=np.int32)

    def do_imf_parity_perm(self, coords, n_dev, n_max, t, perm, ws, ws_bytes):
        n = count(n_dev, n_max)
        C = mat(coords, n, 4, 4, np.int32).astype(np.int64)
        q = np.trunc(C[:, 1:] / t).astype(np.int64)
        cls = (q[:, 0] & 1) | ((q[:, 1] & 1) << 1) | ((q[:, 2] & 1) << 2)
        vec(perm, n, np.int32)[:] = np.argsort(cls, kind="stable").astype(np.int32)

    def do_imf_kernel_map_t_batch(self, jobs, njobs, n_max, cap, K, ld_n):
        offs = offsets(K)
        for j in range(njobs):
            jb = jobs[j]
            n = count(jb.n_out_dev, n_max)
            C = mat(jb.out_coords, max(n, 0), 4, 4, np.int32)
            if jb.perm:
                C = C[vec(jb.perm, n, np.int32)]
            table = self.tables[jb.table_in]
            npad = (n + 127) // 128 * 128
            nbr = mat(jb.nbr_t, K ** 3, npad, ld_n, np.int32)
            nbr[:] = -1
            rows = C.tolist()
            for k, (dx, dy, dz) in enumerate(offs.tolist()):
                col = nbr[k]
                for o, (b, x, y, z) in enumerate(rows):
                    r = table.get((b, x + dx * jb.scale, y + dy * jb.scale, z + dz * jb.scale))
                    if r is not None:
                        col[o] = r
            ntiles = n_max // 128 + 2
            mask = vec(jb.tile_mask, ntiles, np.uint32)
            mask[:] = 0
            for tile in range(npad // 128):
                bits = 0
                for k in range(K ** 3):
                    if (nbr[k, tile * 128:(tile + 1) * 128] >= 0).any():
                        bits |= 1 << k
                mask[tile] = bits

    def do_imf_kernel_map_t(self, out_coords, n_out_dev, n_out_max, table_in, cap, K, scale, nbr_t, ld_n, tile_mask):
        job = _lib.KmapJob(out_coords, n_out_dev, table_in, nbr_t, tile_mask, None, scale, None, None)
        self.do_imf_kernel_map_t_batch([job], 1, n_out_max, cap, K, ld_n)

    def do_imf_batch_segments(self, coords, n_dev, n_max, B, seg):
        n = count(n_dev, n_max)
        b = mat(coords, n, 4, 4, np.int32)[:, 0]
        vec(seg, B + 1, np.int32)[:] = np.searchsorted(b, np.arange(B + 1), side="left").astype(np.int32)

    def do_imf_batch_segments_n(self, coords, n_dev, n_max, B, cap_item, seg, cnt, err):
        n = count(n_dev, n_max)
        b = mat(coords, n, 4, 4, np.int32)[:, 0]
        s = np.searchsorted(b, np.arange(B + 1), side="left").astype(np.int32)
        vec(seg, B + 1, np.int32)[:] = s
        c = np.diff(s)
        vec(cnt, B, np.int32)[:] = np.minimum(c, cap_item)
        e = vec(err, 1, np.int32)
        if (c > cap_item).any():
            e[0] |= 0x20000
        if s[B] != n:
            e[0] |= 0x40000

    # ---- h2 <-> fp32 ---------------------------------------------------------------------------------------------------
    def do_imf_h2_unpack_n(self, H, ldh, n, n_dev, C, KC, X, ldx):
        m = count(n_dev, n)
        mat(X, m, C, ldx)[:] = mat(H, m, C, ldh // 2)

    def do_imf_h2_unpack(self, H, ldh, n, C, KC, X, ldx):
        self.do_imf_h2_unpack_n(H, ldh, n, None, C, KC, X, ldx)

    def do_imf_h2_pack_n(self, X, ldx, n, n_dev, C, KC, H, ldh, err):
        m = count(n_dev, n)
        mat(H, m, C, ldh // 2)[:] = mat(X, m, C, ldx)

    def do_imf_h2_pack(self, X, ldx, n, C, KC, H, ldh, err):
        self.do_imf_h2_pack_n(X, ldx, n, None, C, KC, H, ldh, err)

    def do_imf_h2_unpack_scaled_n(self, H, ldh, n, n_dev, C, KC, mul, X, ldx):
        m = count(n_dev, n)
        mat(X, m, C, ldx)[:] = mat(H, m, C, ldh // 2) * np.float32(mul)

    def do_imf_h2_pack_scaled_n(self, X, ldx, n, n_dev, C, KC, mul, H, ldh, err):
        m = count(n_dev, n)
        mat(H, m, C, ldh // 2)[:] = mat(X, m, C, ldx) * np.float32(mul)

    def do_imf_h2_unpack_l2norm(self, H, ldh, n, n_dev, C, KC, normalize, out_row, Y, ldy):
        m = count(n_dev, n)
        x = mat(H, m, C, ldh // 2).copy()
        if normalize:
            x = x / np.sqrt((x * x).sum(axis=1, keepdims=True))
        rows = vec(out_row, m, np.int32).astype(np.int64) if out_row else np.arange(m)
        mat(Y, int(rows.max()) + 1 if m else 0, C, ldy)[rows] = x

    def do_imf_identity_table(self, n_dev, n_max, nbr_t, ld_n, tile_mask):
        n = count(n_dev, n_max)
        t = vec(nbr_t, ld_n, np.int32)
        t[:] = -1
        t[:n] = np.arange(n, dtype=np.int32)
        vec(tile_mask, ld_n // 128 + 1, np.uint32)[:] = 0
        vec(tile_mask, (n + 127) // 128, np.uint32)[:] = 1

    # ---- convolutions --------------------------------------------------------------------------------------------------
    def do_imf_sparse_conv_h2_pack(self, W, K3, Cin, Cout, kc_in, wmul, packed):
        self.packed[packed] = (vec(W, K3 * Cin * Cout).reshape(K3, Cin, Cout).copy() * np.float32(wmul))

    def do_imf_conv_first_h2_fwd(self, X, ldx, Cin, W, coords, n_dev, n_max, table, cap, K, tstride, Cout, scale, shift, relu, Y, ldy, kc_out):
        n = count(n_dev, n_max)
        C = mat(coords, n, 4, 4, np.int32).tolist()
        x = mat(X, n, Cin, ldx)
        Wk = vec(W, K ** 3 * Cin * Cout).reshape(K ** 3, Cin, Cout)
        t = self.tables[table]
        acc = np.zeros((n, Cout), dtype=np.float32)
        for k, (dx, dy, dz) in enumerate(offsets(K).tolist()):
            idx = np.fromiter((t.get((b, x_ + dx * tstride, y_ + dy * tstride, z_ + dz * tstride), -1) for b, x_, y_, z_ in C),
                              dtype=np.int64, count=n)
            ok = idx >= 0
            if ok.any():
                acc[ok] += x[idx[ok]] @ Wk[k]
        if scale:
            acc = acc * vec(scale, Cout) + vec(shift, Cout)
        if relu:
            acc = np.maximum(acc, 0)
        mat(Y, n, Cout, ldy // 2)[:] = acc

    def do_imf_conv_first_tc_h2_fwd(self, X, ldx, packed, coords, n_dev, n_max, num_items, table, cap, K, Cout, scale, shift, relu, Y, ldy,
                                    kc_out, ws, ws_bytes, err):
        n = count(n_dev, n_max)
        C = mat(coords, n, 4, 4, np.int32).tolist()
        x = mat(X, n, 1, ldx)
        Wk = self.packed[packed]                      # [1, KP, Cout] (already times wmul; scale carries 1 / wmul)
        assert Wk.shape[0] == 1 and Wk.shape[1] >= K ** 3 and Wk.shape[2] == Cout and not Wk[0, K ** 3:].any()
        t = self.tables[table]
        acc = np.zeros((n, Cout), dtype=np.float32)
        for k, (dx, dy, dz) in enumerate(offsets(K).tolist()):
            idx = np.fromiter((t.get((b, x_ + dx, y_ + dy, z_ + dz), -1) for b, x_, y_, z_ in C), dtype=np.int64, count=n)
            ok = idx >= 0
            if ok.any():
                acc[ok] += x[idx[ok]] @ Wk[0, k:k + 1]
        acc = acc * vec(scale, Cout) + vec(shift, Cout)
        if relu:
            acc = np.maximum(acc, 0)
        mat(Y, n, Cout, ldy // 2)[:] = acc

    def do_imf_conv_first_tc_h2_fwd_keep(self, X, ldx, packed, coords, n_dev, n_max, num_items, table, cap, K, Cout, scale, shift, relu, Y, ldy,
                                    kc_out, ws, ws_bytes, err):
        self.do_imf_conv_first_tc_h2_fwd(X, ldx, packed, coords, n_dev, n_max, num_items, table, cap, K, Cout, scale, shift, relu, Y, ldy, kc_out, ws, ws_bytes, err)

    def do_imf_conv_first_tc_release(self, coords, n_dev, n_max, num_items, K, ws, ws_bytes):
        pass          # (the emulated conv1 does not use the grid)

    def do_imf_sparse_conv_g4_fwd_perm(self, X, ldx, kc_in, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, K3, Cin, Cout, scale,
                                       shift, R, ldr, kc_r, relu, Y, ldy, n_y_rows, kc_out, out_row, ws, ws_bytes, err):
        n = count(n_out_dev, n_out_max)
        if n <= 0:
            return
        Wk = self.packed[packed]
        assert Wk.shape == (K3, Cin, Cout)
        nbr = mat(nbr_t, K3, n, ld_n, np.int32)
        # the tile masks must cover every present neighbour (the kernel skips (offset, tile) pairs whose bit is clear)
        masks = vec(tile_mask, (n + 127) // 128, np.uint32)
        for k in range(K3):
            present = np.add.reduceat((nbr[k] >= 0).astype(np.int64), np.arange(0, n, 128)) > 0
            assert not (present & (((masks >> k) & 1) == 0)).any(), "tile mask misses a neighbour"
        n_in = int(nbr.max()) + 1
        x = mat(X, n_in, Cin, ldx // 2)
        acc = np.zeros((n, Cout), dtype=np.float32)
        for k in range(K3):
            idx = nbr[k].astype(np.int64)
            ok = idx >= 0
            if ok.any():
                acc[ok] += x[idx[ok]] @ Wk[k]
        acc = acc * vec(scale, Cout) + vec(shift, Cout)
        if R:
            acc = acc + mat(R, n, Cout, ldr // 2)
        if relu:
            acc = np.maximum(acc, 0)
        if out_row:
            rows = vec(out_row, n, np.int32).astype(np.int64)
            mat(Y, n_y_rows, Cout, ldy // 2)[rows] = acc
        else:
            mat(Y, n, Cout, ldy // 2)[:] = acc

    def do_imf_sparse_conv_g4_fwd(self, X, ldx, kc_in, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, K3, Cin, Cout, scale, shift,
                                  R, ldr, kc_r, relu, Y, ldy, n_y_rows, kc_out, ws, ws_bytes, err):
        self.do_imf_sparse_conv_g4_fwd_perm(X, ldx, kc_in, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, K3, Cin, Cout, scale,
                                            shift, R, ldr, kc_r, relu, Y, ldy, n_y_rows, kc_out, None, ws, ws_bytes, err)

    def do_imf_pointwise_tail_h2_fwd(self, X, ldx, C0, Ca, kca, kcb, W1, C1, W2, b2, C2, n_dev, n_max, normalize, out_row, Y, ldy):
        n = count(n_dev, n_max)
        x = mat(X, n, C0, ldx // 2)
        h = np.maximum(x @ vec(W1, C0 * C1).reshape(C0, C1), 0)
        y = h @ vec(W2, C1 * C2).reshape(C1, C2)
        if b2:
            y = y + vec(b2, C2)
        if normalize:
            y = y / np.linalg.norm(y, axis=1, keepdims=True)
        assert not out_row
        mat(Y, n, C2, ldy)[:] = y

    def do_imf_tail_fused_h2_fwd(self, X, ldx, n_max, n_dev, c0, c1, c2, packed1, scale1, shift1, packed2, scale2, bias2, normalize, out, ldo, err):
        n = count(n_dev, n_max)
        x = mat(X, n, c0, ldx // 2)
        h = (x @ self.packed[packed1][0]) * vec(scale1, c1)
        if shift1:
            h = h + vec(shift1, c1)
        y = (np.maximum(h, 0) @ self.packed[packed2][0]) * vec(scale2, c2)
        if bias2:
            y = y + vec(bias2, c2)
        if normalize:
            y = y / np.sqrt((y * y).sum(axis=1, keepdims=True))
        mat(out, n, c2, ldo)[:] = y

    # ---- image branch --------------------------------------------------------------------------------------------------
    def do_imf_image_conv_table(self, Hin, Win, K, stride, pad, nbr_t, ld_n, tile_mask):
        Hout, Wout = (Hin + 2 * pad - K) // stride + 1, (Win + 2 * pad - K) // stride + 1
        n = Hout * Wout
        tiles = (n + 127) // 128
        nbr = mat(nbr_t, K * K, tiles * 128, ld_n, np.int32)
        nbr[:] = -1
        oy, ox = np.divmod(np.arange(n), Wout)
        for k in range(K * K):
            iy, ix = oy * stride - pad + k // K, ox * stride - pad + k % K
            ok = (iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)
            nbr[k, :n][ok] = (iy * Win + ix)[ok]
        mask = vec(tile_mask, tiles + 1, np.uint32)
        mask[:] = 0
        for tile in range(tiles):
            for k in range(K * K):
                if (nbr[k, tile * 128:(tile + 1) * 128] >= 0).any():
                    mask[tile] |= np.uint32(1 << k)

    def do_imf_image_im2col_h2(self, image, C, H, W, K, stride, pad, Kpad, Y, ldy):
        img = vec(image, C * H * W).reshape(C, H, W)
        Hout, Wout = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
        P = np.zeros((C, H + 2 * pad, W + 2 * pad), dtype=np.float32)
        P[:, pad:pad + H, pad:pad + W] = img
        out = mat(Y, Hout * Wout, Kpad, ldy // 2)
        out[:] = 0
        for ky in range(K):
            for kx in range(K):
                patch = P[:, ky:ky + stride * Hout:stride, kx:kx + stride * Wout:stride]      # [C, Hout, Wout]
                col0 = C * (kx + K * ky)
                out[:, col0:col0 + C] = patch.reshape(C, -1).T

    def do_imf_image_stem_h2_fwd(self, image, H, W, B, packed, scale, shift, ws, ws_bytes, Y, ldy, err):
        Wk = self.packed[packed].reshape(8, 8, 4, 64)                   # [ky (8th = 0)][kx = -1..6][c (4th = 0)][o], already times wmul
        H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        for b in range(B):
            img = vec(image + 4 * b * 3 * H * W, 3 * H * W).reshape(3, H, W)
            P = np.zeros((4, 2 * (H1 - 1) + 7, 2 * (W1 - 1) + 8), dtype=np.float32)      # 3 pad rows on top, 4 pad columns on the left
            P[:3, 3:3 + H, 4:4 + W] = img
            acc = np.zeros((H1 * W1, 64), dtype=np.float32)
            for ky in range(7):
                for kx in range(8):
                    patch = P[:, ky:ky + 2 * H1:2, kx:kx + 2 * W1:2]               # [4, H1, W1]
                    acc += patch.reshape(4, -1).T @ Wk[ky, kx]
            out = np.maximum(acc * vec(scale, 64) + vec(shift, 64), 0)
            mat(Y + 2 * b * H1 * W1 * ldy, H1 * W1, 64, ldy // 2)[:] = out

    # P8 plane layout, emulated as fp32 [B][8 chunks][Hp][Wp][8] in the first half of the buffer (the real one holds hi and lo planes)
    @staticmethod
    def _p8(ptr, H, W, B):
        return vec(ptr, B * 8 * (H + 2) * (W + 2) * 8).reshape(B, 8, H + 2, W + 2, 8)

    def do_imf_image_maxpool_p8(self, X, ldx, kc, Hin, Win, K, stride, pad, Y, B):
        Hout, Wout = (Hin + 2 * pad - K) // stride + 1, (Win + 2 * pad - K) // stride + 1
        out = self._p8(Y, Hout, Wout, B)
        for b in range(B):
            x = mat(X + 2 * b * Hin * Win * ldx, Hin * Win, 64, ldx // 2).reshape(Hin, Win, 64)
            P = np.full((Hin + 2 * pad, Win + 2 * pad, 64), -np.inf, dtype=np.float32)
            P[pad:pad + Hin, pad:pad + Win] = x
            m = np.full((Hout, Wout, 64), -np.inf, dtype=np.float32)
            for ky in range(K):
                for kx in range(K):
                    m = np.maximum(m, P[ky:ky + stride * Hout:stride, kx:kx + stride * Wout:stride])
            out[b, :, 1:-1, 1:-1, :] = m.reshape(Hout, Wout, 8, 8).transpose(2, 0, 1, 3)

    def do_imf_image_conv3x3_p8_fwd(self, X, H, W, B, packed, scale, shift, residual, relu, Y, y_pixel_major, ldy, err):
        Wk = self.packed[packed]                                          # [9, 64, 64], tap kx + 3 ky, already times wmul
        x = self._p8(X, H, W, B)
        assert not x[:, :, 0].any() and not x[:, :, -1].any() and not x[:, :, :, 0].any() and not x[:, :, :, -1].any(), "P8 border must be zero"
        for b in range(B):
            img = x[b].transpose(1, 2, 0, 3).reshape(H + 2, W + 2, 64)      # padded [Hp, Wp, C]
            acc = np.zeros((H * W, 64), dtype=np.float32)
            for ky in range(3):
                for kx in range(3):
                    acc += img[ky:ky + H, kx:kx + W].reshape(H * W, 64) @ Wk[kx + 3 * ky]
            acc = acc * vec(scale, 64) + vec(shift, 64)
            if residual:
                acc = acc + self._p8(residual, H, W, B)[b].transpose(1, 2, 0, 3).reshape(H + 2, W + 2, 64)[1:-1, 1:-1].reshape(H * W, 64)
            if relu:
                acc = np.maximum(acc, 0)
            if y_pixel_major:
                mat(Y + 2 * b * H * W * ldy, H * W, 64, ldy // 2)[:] = acc
            else:
                self._p8(Y, H, W, B)[b, :, 1:-1, 1:-1, :] = acc.reshape(H, W, 8, 8).transpose(2, 0, 1, 3)

    def do_imf_image_maxpool_h2(self, X, ldx, kc, C, Hin, Win, K, stride, pad, Y, ldy):
        Hout, Wout = (Hin + 2 * pad - K) // stride + 1, (Win + 2 * pad - K) // stride + 1
        x = mat(X, Hin * Win, C, ldx // 2).reshape(Hin, Win, C)
        P = np.full((Hin + 2 * pad, Win + 2 * pad, C), -np.inf, dtype=np.float32)
        P[pad:pad + Hin, pad:pad + Win] = x
        m = np.full((Hout, Wout, C), -np.inf, dtype=np.float32)
        for ky in range(K):
            for kx in range(K):
                m = np.maximum(m, P[ky:ky + stride * Hout:stride, kx:kx + stride * Wout:stride])
        mat(Y, Hout * Wout, C, ldy // 2)[:] = m.reshape(-1, C)

    def do_imf_image_im2col_h2_batch(self, image, C, H, W, K, stride, pad, Kpad, Y, ldy, B):
        Hout, Wout = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
        for b in range(B):
            self.do_imf_image_im2col_h2(image + 4 * b * C * H * W, C, H, W, K, stride, pad, Kpad, Y + 2 * b * Hout * Wout * ldy, ldy)

    def do_imf_image_maxpool_h2_batch(self, X, ldx, kc, C, Hin, Win, K, stride, pad, Y, ldy, B):
        Hout, Wout = (Hin + 2 * pad - K) // stride + 1, (Win + 2 * pad - K) // stride + 1
        for b in range(B):
            self.do_imf_image_maxpool_h2(X + 2 * b * Hin * Win * ldx, ldx, kc, C, Hin, Win, K, stride, pad, Y + 2 * b * Hout * Wout * ldy, ldy)

    # ---- attention fusion ----------------------------------------------------------------------------------------------
    def do_imf_attention_kv(self, w, tokens, L, channel_major, kv, ws, ws_bytes):
        assert not channel_major
        t = mat(tokens, L, w.dim, w.dim)
        cn = layer_norm(t, vec(w.ln_c_w, w.dim), vec(w.ln_c_b, w.dim))
        p = cn @ vec(w.wkv, 2 * w.inner * w.dim).reshape(2 * w.inner, w.dim).T
        self.kv[kv] = (p[:, :w.inner].copy(), p[:, w.inner:].copy())

    def do_imf_attention_fusion_fwd_m(self, w, P, ldp, M, m_dev, kv, L, out, ldo, ws, ws_bytes):
        m = count(m_dev, M)
        if m <= 0:
            return
        lat, inner = w.latent, w.inner
        x = mat(P, m, lat, ldp).copy()
        K, V = self.kv[kv]
        assert K.shape[0] == L
        q = layer_norm(x, vec(w.ln_q_w, lat), vec(w.ln_q_b, lat)) @ vec(w.wq, inner * lat).reshape(inner, lat).T
        s = (q @ K.T) * np.float32(inner ** -0.5)
        s = np.exp(s - s.max(axis=1, keepdims=True))
        a = s / s.sum(axis=1, keepdims=True)
        x1 = (a @ V) @ vec(w.wo, lat * inner).reshape(lat, inner).T + vec(w.bo, lat) + x
        h = layer_norm(x1, vec(w.ln_f_w, lat), vec(w.ln_f_b, lat)) @ vec(w.w1, 8 * lat * lat).reshape(8 * lat, lat).T + vec(w.b1, 8 * lat)
        h = h[:, :4 * lat] * gelu(np.ascontiguousarray(h[:, 4 * lat:]))
        mat(out, m, lat, ldo)[:] = h @ vec(w.w2, lat * 4 * lat).reshape(lat, 4 * lat).T + vec(w.b2, lat) + x1

    def do_imf_attention_kv_batched(self, w, wp, tokens, L, B, kv, ws, ws_bytes, err):
        t = mat(tokens, B * L, w.dim, w.dim)
        cn = layer_norm(t, vec(w.ln_c_w, w.dim), vec(w.ln_c_b, w.dim))
        p = cn @ vec(w.wkv, 2 * w.inner * w.dim).reshape(2 * w.inner, w.dim).T
        self.kv[kv] = [(p[b * L:(b + 1) * L, :w.inner].copy(), p[b * L:(b + 1) * L, w.inner:].copy()) for b in range(B)]

    def do_imf_attention_fusion_fwd_batched(self, w, wp, P, ldp, M, m_dev, seg_dev, cnt_dev, B, kv, L, out, ldo, ws, ws_bytes, err):
        seg, cnt = vec(seg_dev, B, np.int32), vec(cnt_dev, B, np.int32)
        m = count(m_dev, M)
        saved = self.kv[kv]
        try:
            for b in range(B):          # per item: its rows against its image (model/resunet.py:240-271); never past the M rows
                lo, hi = int(seg[b]), min(int(seg[b]) + int(cnt[b]), m)
                if hi <= lo:
                    continue
                self.kv[kv] = saved[b]
                self.do_imf_attention_fusion_fwd_m(w, P + 4 * ldp * lo, ldp, hi - lo, None, kv, L, out + 4 * ldo * lo, ldo, ws, ws_bytes)
        finally:
            self.kv[kv] = saved

    def do_imf_attention_fusion_fwd(self, w, P, ldp, M, kv, L, out, ldo, ws, ws_bytes):
        self.do_imf_attention_fusion_fwd_m(w, P, ldp, M, None, kv, L, out, ldo, ws, ws_bytes)

    def do_imf_transpose_tokens(self, X, L, C, Y):
        mat(Y, C, L, L)[:] = mat(X, L, C, C).T
