"""GPU: sparse convolution kernels vs the oracle (fp32, tolerance 1e-4 relative as the north star states)."""
import numpy as np
import pytest
import torch

from imfnet_b200 import _lib, synthetic
from oracle import sparse_ops

pytestmark = pytest.mark.gpu
RTOL = 1e-4      # north-star tolerance; observed error is ~1e-6 (fp32 accumulation-order noise)
# tensor-core tiers: the operands carry 22 mantissa bits, but the MMA unit adds into its fp32 accumulator with truncation, so the
# error grows with the number of accumulation steps (27 offsets x Cin/16 x 3 products): observed up to 8e-6 at Cin = 128-256
H2_RTOL = 3e-5


def close(a, b, rtol=RTOL):
    scale = float(b.abs().max()) + 1e-30
    err = float((a - b).abs().max()) / scale
    assert err < rtol, f"max abs err / max|ref| = {err:.3e}"


@pytest.fixture(scope="module")
def frag():
    coords, _ = synthetic.make_fragment(6000, 0.05, seed=4)
    ocm = sparse_ops.CoordinateManager(coords)
    for t in (1, 2, 4):
        ocm.stride(t, 2)
    from imfnet_b200.sparse import CoordinateManager
    cm = CoordinateManager(torch.from_numpy(coords).cuda())
    cm.build_pyramid([2, 4, 8])
    return coords, ocm, cm


def run_conv(X, W, nbr, n_out, scale=None, shift=None, R=None, relu=False):
    L = _lib.lib()
    Y = torch.full((n_out, W.shape[-1]), float("nan"), device="cuda")
    _lib.check(L.imf_sparse_conv_fwd(X.data_ptr(), X.stride(0), W.data_ptr(), nbr.data_ptr(), None, n_out, W.shape[0],
                                     W.shape[1], W.shape[2], _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(R),
                                     0 if R is None else R.stride(0), int(relu), Y.data_ptr(), Y.stride(0), _lib.cur_stream()))
    torch.cuda.synchronize()
    return Y.cpu()


@pytest.mark.parametrize("t_in,t_out,tr,cin,cout", [
    (1, 1, False, 32, 32), (1, 1, False, 64, 64), (1, 2, False, 32, 64), (2, 2, False, 64, 64), (2, 4, False, 64, 128),
    (4, 4, False, 128, 128), (4, 8, False, 128, 256), (8, 8, False, 256, 256), (8, 4, True, 256, 128), (4, 2, True, 256, 64),
    (2, 1, True, 128, 64)])
def test_conv_matches_oracle(frag, t_in, t_out, tr, cin, cout):
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(cin * 1000 + cout + t_in)
    n_in, n_out = len(ocm.get(t_in)), len(ocm.get(t_out))
    X = torch.randn(n_in, cin, generator=g)
    W = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    onbr = ocm.table(t_in, t_out, 3, tr)
    ref = sparse_ops.conv_forward(X, W, onbr)
    nbr = cm.table(t_in, t_out, 3, tr)
    assert np.array_equal(nbr.cpu().numpy(), onbr)
    close(run_conv(X.cuda(), W.cuda(), nbr, n_out), ref)


def h2_pack(X, kc, ld_extra=0):
    L = _lib.lib()
    n, C = X.shape
    H = torch.zeros((n, 2 * C + ld_extra), dtype=torch.float16, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.imf_h2_pack(X.data_ptr(), X.stride(0), n, C, kc, H.data_ptr(), H.stride(0), err.data_ptr(), _lib.cur_stream()))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return H


def h2_unpack(H, C, kc):
    L = _lib.lib()
    n = H.shape[0]
    X = torch.empty((n, C), dtype=torch.float32, device="cuda")
    _lib.check(L.imf_h2_unpack(H.data_ptr(), H.stride(0), n, C, kc, X.data_ptr(), C, _lib.cur_stream()))
    torch.cuda.synchronize()
    return X


def test_h2_pack_unpack_roundtrip_keeps_22_bits():
    g = torch.Generator().manual_seed(3)
    X = (torch.randn(1000, 96, generator=g) * torch.logspace(-3, 3, 96)).cuda()
    for kc in (32,):
        back = h2_unpack(h2_pack(X, kc), 96, kc)
        assert bool(((back - X).abs() <= torch.maximum(X.abs() * 2.0 ** -21, torch.tensor(6.1e-8, device="cuda"))).all())
    X = torch.randn(777, 128, generator=g).cuda()
    back = h2_unpack(h2_pack(X, 64, ld_extra=24), 128, 64)
    assert bool(((back - X).abs() <= torch.maximum(X.abs() * 2.0 ** -21, torch.tensor(6.1e-8, device="cuda"))).all())


def run_conv_g4(X, W, tab, n_out, scale, shift, R=None, relu=False, split=True, kc_out=None, n_dev=None, extra_rows=0, kc_r=None):
    """fp32 in/out wrapper of the TMA-gather kernel: pack -> conv -> unpack.  tab = CoordinateManager.table_t(...)."""
    L = _lib.lib()
    nbr_t, ld_n, tile_mask = tab
    K3, cin, cout = W.shape
    kc_in = 64 if cin % 64 == 0 else 32
    kc_out = kc_out or (64 if cout % 64 == 0 else 32)
    wmul = 2.0 ** np.floor(np.log2(2048.0 / float(W.abs().max())))
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(K3, cin, cout, kc_in)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), K3, cin, cout, kc_in, wmul, packed.data_ptr(), _lib.cur_stream()))
    Xh = h2_pack(X, kc_in, ld_extra=8)
    kc_r = kc_r or kc_out
    Rh = None if R is None else h2_pack(R, kc_r, ld_extra=16)
    Yh = torch.full((n_out + extra_rows, 2 * cout + 8), float("nan"), dtype=torch.float16, device="cuda")
    ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(cout)) if split else 0
    ws = torch.zeros(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    sc = (scale / wmul).contiguous()
    _lib.check(L.imf_sparse_conv_g4_fwd(Xh.data_ptr(), Xh.stride(0), kc_in, packed.data_ptr(), nbr_t.data_ptr(), ld_n,
                                        tile_mask.data_ptr(), _lib.ptr(n_dev), n_out, K3, cin, cout, sc.data_ptr(), shift.data_ptr(),
                                        _lib.ptr(Rh), 0 if Rh is None else Rh.stride(0), kc_r, int(relu), Yh.data_ptr(), Yh.stride(0),
                                        Yh.shape[0], kc_out, ws.data_ptr() if split else None, ws_bytes, err.data_ptr(),
                                        _lib.cur_stream()))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return h2_unpack(Yh, cout, kc_out).cpu()


@pytest.mark.parametrize("t_in,t_out,tr", [(1, 1, False), (1, 2, False), (4, 8, False), (8, 4, True), (2, 1, True)])
def test_offset_major_table_matches_oracle(frag, t_in, t_out, tr):
    """imf_kernel_map_t = the oracle's neighbour table transposed, -1 padding up to the tile boundary, exact per-tile offset masks."""
    coords, ocm, cm = frag
    onbr = ocm.table(t_in, t_out, 3, tr)                      # [n_out, 27]
    nbr_t, ld_n, tile_mask = cm.table_t(t_in, t_out, 3, tr)
    n = onbr.shape[0]
    got = nbr_t.cpu().numpy()
    assert np.array_equal(got[:, :n], onbr.T)
    assert np.all(got[:, n:(n + 127) // 128 * 128] == -1)
    tm = tile_mask.cpu().numpy().astype(np.uint32)
    for t in range((n + 127) // 128):
        blk = onbr[t * 128:(t + 1) * 128] >= 0
        exp = sum(1 << k for k in range(27) if blk[:, k].any())
        assert int(tm[t]) == exp


@pytest.mark.parametrize("t_in,t_out,tr,cin,cout", [
    (1, 1, False, 32, 32), (1, 1, False, 64, 64), (1, 2, False, 32, 64), (2, 2, False, 64, 64), (2, 4, False, 64, 128),
    (4, 4, False, 128, 128), (4, 8, False, 128, 256), (8, 8, False, 256, 256), (8, 4, True, 256, 128), (4, 2, True, 256, 64),
    (2, 1, True, 128, 64)])
@pytest.mark.parametrize("split", [False, True])
def test_g4_persistent_conv_matches_oracle(frag, t_in, t_out, tr, cin, cout, split):
    """Persistent tcgen05 kind::f16 kernel (offset-major table, shared weight slabs, TMA-store epilogue): same contract and tolerance
    as the h2 kernel."""
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(cin * 1000 + cout + t_in)
    n_in, n_out = len(ocm.get(t_in)), len(ocm.get(t_out))
    X = torch.randn(n_in, cin, generator=g)
    W = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    R = torch.randn(n_out, cout, generator=g)
    ref = torch.relu(sparse_ops.conv_forward(X, W, ocm.table(t_in, t_out, 3, tr)) * scale + shift + R)
    out = run_conv_g4(X.cuda(), W.cuda(), cm.table_t(t_in, t_out, 3, tr), n_out, scale.cuda(), shift.cuda(), R.cuda(), True, split)
    close(out, ref, H2_RTOL)


@pytest.mark.parametrize("n,cin,cout", [(50000, 64, 64), (50000, 32, 32), (90000, 32, 128), (19200, 64, 64), (14107, 64, 64), (333, 64, 64)])
def test_g4_conv_equals_simt_conv_large(n, cin, cout):
    """Row-mode partition (several sub-tiles per CTA, several passes when the accumulators exceed TMEM) against the fp32 SIMT kernel."""
    coords, _ = synthetic.make_fragment(n, 0.025, 0)
    from imfnet_b200.sparse import CoordinateManager
    cm = CoordinateManager(torch.from_numpy(coords).cuda())
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, cin, device="cuda", generator=g)
    W = torch.randn(27, cin, cout, device="cuda", generator=g) / 40
    a = run_conv(X, W, cm.table(1, 1, 3, False), n)
    one, zero = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    b = run_conv_g4(X, W, cm.table_t(1, 1, 3, False), n, one, zero, split=True)
    close(b, a, H2_RTOL)


def test_g4_conv_device_side_row_count(frag):
    """n_out_dev < n_out_max: only the first n rows are computed (their neighbours all lie below n too), rows >= roundup32(n) stay untouched."""
    coords, ocm, cm = frag
    from imfnet_b200.sparse import CoordinateManager
    n_max = len(coords)
    n = 2500
    sub = CoordinateManager(torch.from_numpy(coords[:n].copy()).cuda())
    L = _lib.lib()
    # table built with the device-side count on the full-size allocation
    ld_n = (n_max + 127) // 128 * 128
    nbr_t = torch.full((27, ld_n), 12345, dtype=torch.int32, device="cuda")
    tile_mask = torch.empty(ld_n // 128 + 1, dtype=torch.int32, device="cuda")
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    lvl = sub.level(1)
    full = torch.from_numpy(coords).cuda()
    _lib.check(L.imf_kernel_map_t(full.data_ptr(), n_dev.data_ptr(), n_max, lvl.table.data_ptr(), lvl.capacity, 3, 1, nbr_t.data_ptr(), ld_n,
                                  tile_mask.data_ptr(), _lib.cur_stream()))
    g = torch.Generator().manual_seed(5)
    X = torch.randn(n_max, 64, generator=g)
    W = torch.randn(27, 64, 64, generator=g) / 40
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    ocm_sub = sparse_ops.CoordinateManager(coords[:n])
    ref = torch.relu(sparse_ops.conv_forward(X[:n], W, ocm_sub.table(1, 1, 3, False)) * scale + shift)
    for split in (False, True):
        out = run_conv_g4(X.cuda(), W.cuda(), (nbr_t, ld_n, tile_mask), n_max, scale.cuda(), shift.cuda(), None, True, split, n_dev=n_dev)
        close(out[:n], ref, H2_RTOL)
        assert bool(torch.isnan(out[(n + 31) // 32 * 32:]).all())


def test_h2_first_conv_and_tail(frag):
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(11)
    n = len(coords)
    L = _lib.lib()
    X = torch.rand(n, 1, generator=g) + 0.5
    W = torch.randn(125, 1, 32, generator=g) / 11
    scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    ref = sparse_ops.conv_forward(X, W, ocm.table(1, 1, 5, False)) * scale + shift
    lvl = cm.level(1)
    Yh = torch.zeros(n, 64, dtype=torch.float16, device="cuda")
    X_d, W_d, sc_d, sh_d = X.cuda(), W.cuda(), scale.cuda(), shift.cuda()
    _lib.check(L.imf_conv_first_h2_fwd(X_d.data_ptr(), 1, 1, W_d.data_ptr(), lvl.coords.data_ptr(), None, n, lvl.table.data_ptr(),
                                       lvl.capacity, 5, 1, 32, sc_d.data_ptr(), sh_d.data_ptr(), 0, Yh.data_ptr(), 64, 32,
                                       _lib.cur_stream()))
    close(h2_unpack(Yh, 32, 32).cpu(), ref, H2_RTOL)
    # tail on a two-section h2 matrix [64 | 32], scattered through out_row
    Xt = torch.randn(n, 96, generator=g)
    W1, W2, b2 = torch.randn(96, 64, generator=g) / 10, torch.randn(64, 32, generator=g) / 8, torch.randn(32, generator=g)
    ref = torch.relu(Xt @ W1) @ W2 + b2
    ref = ref / torch.norm(ref, p=2, dim=1, keepdim=True)
    Ha, Hb = h2_pack(Xt[:, :64].contiguous().cuda(), 64), h2_pack(Xt[:, 64:].contiguous().cuda(), 32)
    H = torch.cat([Ha, Hb], dim=1).contiguous()
    perm = torch.randperm(n, generator=g).to(torch.int32)
    Y = torch.empty(n, 32, device="cuda")
    W1_d, W2_d, b2_d, perm_d = W1.cuda(), W2.cuda(), b2.cuda(), perm.cuda()
    _lib.check(L.imf_pointwise_tail_h2_fwd(H.data_ptr(), 192, 96, 64, 64, 32, W1_d.data_ptr(), 64, W2_d.data_ptr(), b2_d.data_ptr(),
                                           32, None, n, 1, perm_d.data_ptr(), Y.data_ptr(), 32, _lib.cur_stream()))
    torch.cuda.synchronize()
    out = torch.empty_like(ref)
    out[:] = Y.cpu()
    close(out[perm.long()], ref, H2_RTOL)


def test_conv_fused_epilogue_and_strided_operands(frag):
    """BatchNorm affine + residual + ReLU, reading a column window of a wider buffer and writing into another."""
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(7)
    n = len(coords)
    wide = torch.randn(n, 96, generator=g)
    W = torch.randn(27, 32, 64, generator=g) / 30
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    R = torch.randn(n, 80, generator=g)
    onbr = ocm.table(1, 1, 3, False)
    ref = torch.relu(sparse_ops.conv_forward(wide[:, 64:96].contiguous(), W, onbr) * scale + shift + R[:, 16:80])
    wide_d, R_d, W_d, sc_d, sh_d = wide.cuda(), R.cuda(), W.cuda(), scale.cuda(), shift.cuda()   # keep device copies alive
    out = torch.zeros(n, 128, device="cuda")
    L = _lib.lib()
    _lib.check(L.imf_sparse_conv_fwd(wide_d.data_ptr() + 64 * 4, 96, W_d.data_ptr(), cm.table(1, 1, 3, False).data_ptr(),
                                     None, n, 27, 32, 64, sc_d.data_ptr(), sh_d.data_ptr(),
                                     R_d.data_ptr() + 16 * 4, 80, 1, out.data_ptr() + 64 * 4, 128, _lib.cur_stream()))
    torch.cuda.synchronize()
    close(out[:, 64:].cpu(), ref)
    assert float(out[:, :64].abs().max()) == 0.0          # neighbouring columns untouched


@pytest.mark.parametrize("cin,cout,K", [(1, 32, 5), (3, 32, 5), (1, 32, 3), (1, 64, 5)])
def test_first_conv_matches_oracle(frag, cin, cout, K):
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(cin + cout + K)
    n = len(coords)
    X = torch.rand(n, cin, generator=g) + 0.5
    W = torch.randn(K ** 3, cin, cout, generator=g) / np.sqrt(K ** 3 * cin)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = sparse_ops.conv_forward(X, W, ocm.table(1, 1, K, False)) * scale + shift
    L = _lib.lib()
    lvl = cm.level(1)
    Y = torch.empty(n, cout, device="cuda")
    X_d, W_d, sc_d, sh_d = X.cuda(), W.cuda(), scale.cuda(), shift.cuda()      # keep device copies alive across the call
    _lib.check(L.imf_conv_first_fwd(X_d.data_ptr(), cin, cin, W_d.data_ptr(), lvl.coords.data_ptr(), None, n,
                                    lvl.table.data_ptr(), lvl.capacity, K, 1, cout, sc_d.data_ptr(),
                                    sh_d.data_ptr(), 0, Y.data_ptr(), cout, _lib.cur_stream()))
    torch.cuda.synchronize()
    close(Y.cpu(), ref)


def run_conv_first_tc(coords_np, X, W, scale, shift, K, num_items, n_dev=None):
    """imf_conv_first_tc_h2_fwd (dense-grid / hash-probe neighbour expansion + one-offset tensor-core convolution), fp32 in/out."""
    from imfnet_b200.sparse import CoordinateManager
    L = _lib.lib()
    n, cout = len(coords_np), W.shape[2]
    cm = CoordinateManager(torch.from_numpy(coords_np).cuda(), check=False)
    lvl = cm.level(1)
    KP = int(L.imf_conv_first_tc_columns(K))
    w1 = torch.zeros((1, KP, cout), device="cuda")
    w1[0, :K ** 3] = W[:, 0, :].cuda()
    wmul = 2.0 ** np.floor(np.log2(2048.0 / float(w1.abs().max())))
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, KP, cout, 64)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(w1.data_ptr(), 1, KP, cout, 64, wmul, packed.data_ptr(), _lib.cur_stream()))
    ws_bytes = int(L.imf_conv_first_tc_workspace_bytes(n, K))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    kco = 64 if cout % 64 == 0 else 32
    Yh = torch.full((n, 2 * cout), float("nan"), dtype=torch.float16, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    X_d, sc_d, sh_d = X.cuda().contiguous(), (scale / wmul).cuda().contiguous(), shift.cuda().contiguous()
    _lib.check(L.imf_conv_first_tc_h2_fwd(X_d.data_ptr(), X_d.stride(0), packed.data_ptr(), lvl.coords.data_ptr(), _lib.ptr(n_dev), n, num_items,
                                          lvl.table.data_ptr(), lvl.capacity, K, cout, sc_d.data_ptr(), sh_d.data_ptr(), 0, Yh.data_ptr(),
                                          2 * cout, kco, ws.data_ptr(), ws_bytes, err.data_ptr(), _lib.cur_stream()))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    use_grid = int(ws[:4].view(torch.int32).item())
    return h2_unpack(Yh, cout, kco).cpu(), use_grid


@pytest.mark.parametrize("cout,K", [(32, 5), (32, 3), (64, 5), (32, 1)])
def test_first_conv_tensor_core_path_matches_oracle(frag, cout, K):
    """conv1 with one input channel as neighbour expansion (dense row-index grid) + tcgen05 product, against the oracle's convolution."""
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(cout + K)
    n = len(coords)
    X = torch.rand(n, 1, generator=g) + 0.5
    W = torch.randn(K ** 3, 1, cout, generator=g) / np.sqrt(K ** 3)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = sparse_ops.conv_forward(X, W, ocm.table(1, 1, K, False)) * scale + shift
    out, use_grid = run_conv_first_tc(coords, X, W, scale, shift, K, 1)
    assert use_grid == 1
    close(out, ref, H2_RTOL)


def test_first_conv_tensor_core_path_batches_fallback_and_device_count():
    """(a) three batch items with different bounding boxes (one of them empty) share the grid; (b) scattered voxels whose boxes exceed
    the grid budget take the hash-probe fallback; (c) a foreign batch index falls back per voxel; (d) a device-side row count."""
    g = torch.Generator().manual_seed(9)
    rng = np.random.default_rng(9)
    K, cout = 5, 32
    W = torch.randn(125, 1, cout, generator=g) / 11
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1

    def check(coords, num_items, expect_grid, n_dev=None, n_eff=None):
        n = len(coords)
        X = torch.rand(n, 1, generator=g) + 0.5
        m = n if n_eff is None else n_eff
        ocm = sparse_ops.CoordinateManager(coords[:m])
        ref = sparse_ops.conv_forward(X[:m], W, ocm.table(1, 1, K, False)) * scale + shift
        out, use_grid = run_conv_first_tc(coords, X, W, scale, shift, K, num_items, n_dev)
        assert use_grid == expect_grid
        close(out[:m], ref, H2_RTOL)

    a, _ = synthetic.make_fragment(3000, 0.05, seed=1)
    b, _ = synthetic.make_fragment(2000, 0.05, seed=2)
    b = b.copy(); b[:, 0] = 2; b[:, 1:] += np.array([500, -300, 40], dtype=np.int32)
    check(np.concatenate([a, b]), 3, 1)                                             # item 1 is empty
    centres = rng.integers(-20000, 20000, size=(60, 1, 3))
    far = (centres + rng.integers(-3, 4, size=(60, 50, 3))).reshape(-1, 3).astype(np.int32)          # clusters far apart: neighbours exist
    far = far[np.sort(np.unique(far, axis=0, return_index=True)[1])]
    far = np.concatenate([np.zeros((len(far), 1), np.int32), far], axis=1)
    check(far, 1, 0)
    c = a.copy(); c[-100:, 0] = 7                                                   # batch index outside num_items for the last rows
    check(c, 1, 1)
    n_dev = torch.tensor([2100], dtype=torch.int32, device="cuda")
    check(a, 1, 1, n_dev=n_dev, n_eff=2100)


def test_neighbour_tables_from_the_dense_grid_equal_the_hash_tables():
    """imf_conv_first_tc_h2_fwd_keep leaves the stride-1 row-index grid populated; imf_kernel_map_t_batch with dense_meta / dense_cells
    must give bit-identical tables and tile masks to the hash probes -- same-level (1 -> 1) and strided (1 -> 2) maps, two batch items,
    a row with a foreign batch index (hash fallback) -- and imf_conv_first_tc_release must leave the grid all-empty again."""
    import ctypes as C
    from imfnet_b200.sparse import CoordinateManager
    a, _ = synthetic.make_fragment(5000, 0.05, seed=11)
    b, _ = synthetic.make_fragment(3000, 0.05, seed=12)
    b = b.copy(); b[:, 0] = 1; b[:, 1:] += np.array([40, -25, 7], dtype=np.int32)
    coords = np.concatenate([a, b])
    coords[-3:, 0] = 5                                           # foreign batch index (num_items = 2)
    n = len(coords)
    L = _lib.lib()
    s = _lib.cur_stream()
    cm = CoordinateManager(torch.from_numpy(coords).cuda(), check=False)
    cm.build_pyramid([2])
    l1, l2 = cm.level(1), cm.level(2)
    K, cout = 5, 32
    g = torch.Generator().manual_seed(1)
    W = torch.randn(125, 1, cout, generator=g) / 11
    KP = int(L.imf_conv_first_tc_columns(K))
    w1 = torch.zeros((1, KP, cout), device="cuda")
    w1[0, :125] = W[:, 0, :].cuda()
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, KP, cout, 64)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(w1.data_ptr(), 1, KP, cout, 64, 1024.0, packed.data_ptr(), s))
    ws_bytes = int(L.imf_conv_first_tc_workspace_bytes(n, K))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")          # zero once, as the plans do
    X = torch.ones(n, 1, device="cuda")
    sc, sh = torch.full((cout,), 1.0 / 1024, device="cuda"), torch.zeros(cout, device="cuda")
    Yh = torch.zeros((n, 2 * cout), dtype=torch.float16, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    gm, gc = C.c_void_p(), C.c_void_p()
    _lib.check(L.imf_conv_first_tc_grid(ws.data_ptr(), n, K, C.byref(gm), C.byref(gc)))
    ld_n = (n + 127) // 128 * 128

    def tables(dense):
        out = []
        jobs = (_lib.KmapJob * 2)()
        for i, (dst, scale) in enumerate(((l1, 1), (l2, 1))):
            nbr_t = torch.full((27, ld_n), -9, dtype=torch.int32, device="cuda")
            mask = torch.full((ld_n // 128 + 2,), -9, dtype=torch.int32, device="cuda")
            jobs[i] = _lib.KmapJob(dst.coords.data_ptr(), None if i == 0 else None, l1.table.data_ptr(), nbr_t.data_ptr(), mask.data_ptr(), None, scale,
                                   gm.value if dense else None, gc.value if dense else None)
            out.append((nbr_t, mask, dst.n))
        # (both jobs share n_out_max = n; the stride-2 level has fewer rows: its count is passed on the device)
        n2 = torch.tensor([l2.n], dtype=torch.int32, device="cuda")
        jobs[1].n_out_dev = n2.data_ptr()
        _lib.check(L.imf_kernel_map_t_batch(jobs, 2, n, l1.capacity, 3, ld_n, s))
        torch.cuda.synchronize()
        return out

    for rep in range(2):                                         # twice: the second use relies on the release of the first
        _lib.check(L.imf_conv_first_tc_h2_fwd_keep(X.data_ptr(), 1, packed.data_ptr(), l1.coords.data_ptr(), None, n, 2, l1.table.data_ptr(), l1.capacity,
                                                   K, cout, sc.data_ptr(), sh.data_ptr(), 0, Yh.data_ptr(), 2 * cout, 32, ws.data_ptr(), ws_bytes,
                                                   err.data_ptr(), s))
        torch.cuda.synchronize()
        assert int(ws[:4].view(torch.int32).item()) == 1, "the two boxes must fit the grid budget"
        dense, plain = tables(True), tables(False)
        for (nd, md, cnt), (nh, mh, _c) in zip(dense, plain):
            npad = (cnt + 127) // 128 * 128
            assert torch.equal(nd[:, :npad], nh[:, :npad]) and torch.equal(md[: npad // 128 + 1], mh[: npad // 128 + 1])
            assert int((nh[:, :cnt] >= 0).sum()) > 5 * cnt          # (a real table: ~12 neighbours per row)
        _lib.check(L.imf_conv_first_tc_release(l1.coords.data_ptr(), None, n, 2, K, ws.data_ptr(), ws_bytes, s))
        torch.cuda.synchronize()
        off = gc.value - ws.data_ptr()
        cells = ws[off:off + 4 * (512 * n)].view(torch.int32)
        feats = ws[gm.value - ws.data_ptr() + 14592:off].view(torch.int32)          # the feature grid lies between the header and the row grid
        assert int(cells.abs().max()) == 0 and int(feats.abs().max()) == 0, "release must leave both grids empty"
    assert int(err.item()) == 0


@pytest.mark.parametrize("normalize", [True, False])
def test_pointwise_tail_matches_oracle(normalize):
    g = torch.Generator().manual_seed(5)
    n = 3001
    X = torch.randn(n, 96, generator=g)
    W1, W2, b2 = torch.randn(96, 64, generator=g) / 10, torch.randn(64, 32, generator=g) / 8, torch.randn(32, generator=g)
    ref = torch.relu(X @ W1) @ W2 + b2
    if normalize:
        ref = ref / torch.norm(ref, p=2, dim=1, keepdim=True)
    L = _lib.lib()
    Y = torch.empty(n, 32, device="cuda")
    X_d, W1_d, W2_d, b2_d = X.cuda(), W1.cuda(), W2.cuda(), b2.cuda()          # keep device copies alive across the call
    _lib.check(L.imf_pointwise_tail_fwd(X_d.data_ptr(), 96, 96, W1_d.data_ptr(), 64, W2_d.data_ptr(),
                                        b2_d.data_ptr(), 32, None, n, int(normalize), Y.data_ptr(), 32, _lib.cur_stream()))
    torch.cuda.synchronize()
    close(Y.cpu(), ref)


@pytest.mark.parametrize("n,c0,normalize,n_eff", [(1, 96, True, None), (127, 96, True, None), (129, 64, True, None), (3001, 96, False, None),
                                                  (40000, 96, True, None), (40000, 32, True, 38017), (150 * 128 * 2 + 5, 96, True, None)])
def test_fused_tail_kernel_matches_fp32(n, c0, normalize, n_eff):
    """csrc/tail_fused.cu (conv1_tr -> ReLU -> final + bias -> L2 norm in one tcgen05 kernel) against the fp32 formula of
    model/resunet.py:216-233, at tile-boundary sizes, more tiles than two per SM, a device-side row count and un-normalised output."""
    g = torch.Generator().manual_seed(n + c0)
    X = torch.randn(n, c0, generator=g) * 3
    W1, W2, b2 = torch.randn(c0, 64, generator=g) / 10, torch.randn(64, 32, generator=g) / 8, torch.randn(32, generator=g)
    sh1 = torch.randn(64, generator=g) * 0.1
    m = n if n_eff is None else n_eff
    ref = torch.relu(X[:m].double() @ W1.double() + sh1.double()) @ W2.double() + b2.double()
    if normalize:
        ref = ref / torch.norm(ref, p=2, dim=1, keepdim=True)
    L = _lib.lib()
    s = _lib.cur_stream()
    Xd = X.cuda()
    H = torch.zeros(n, c0, device="cuda")                       # h2 footprint of [n, c0] = fp32 footprint
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.imf_h2_pack(Xd.data_ptr(), c0, n, c0, 32, H.data_ptr(), 2 * c0, err.data_ptr(), s))
    packs = []
    for W, cin, cout, kci in ((W1, c0, 64, 32), (W2, 64, 32, 64)):
        Wd = W.reshape(1, cin, cout).contiguous().cuda()
        wmul = 2.0 ** np.floor(np.log2(2048.0 / float(W.abs().max())))
        buf = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, cin, cout, kci)), dtype=torch.uint8, device="cuda")
        _lib.check(L.imf_sparse_conv_h2_pack(Wd.data_ptr(), 1, cin, cout, kci, float(wmul), buf.data_ptr(), s))
        packs.append((buf, torch.full((cout,), 1.0 / wmul, device="cuda"), Wd))
    out = torch.full((n, 32), float("nan"), device="cuda")
    n_dev = None if n_eff is None else torch.tensor([n_eff], dtype=torch.int32, device="cuda")
    sh1_d, b2_d = sh1.cuda(), b2.cuda()
    for _ in range(2):          # twice: the second launch must not depend on anything the first left behind
        _lib.check(L.imf_tail_fused_h2_fwd(H.data_ptr(), 2 * c0, n, _lib.ptr(n_dev), c0, 64, 32, packs[0][0].data_ptr(), packs[0][1].data_ptr(),
                                           sh1_d.data_ptr(), packs[1][0].data_ptr(), packs[1][1].data_ptr(), b2_d.data_ptr(), int(normalize),
                                           out.data_ptr(), 32, err.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    close(out[:m].cpu().double(), ref, H2_RTOL)
    tile_end = min(n, (m + 127) // 128 * 128)
    assert torch.all(out[m:tile_end] == 0) and torch.isnan(out[tile_end:]).all()          # rows past the count: zeros inside the last tile, untouched beyond


@pytest.mark.parametrize("H,W,B", [(48, 64, 1), (37, 53, 3), (120, 160, 2), (480, 640, 2)])
def test_fused_stem_kernel_matches_conv2d(H, W, B):
    """csrc/stem_fused.cu (7x7 / 2 / 3 convolution + folded BatchNorm + ReLU as an implicit GEMM on a pre-split image) against
    torch's conv2d in float64 (model/resnet.py:195-207), odd sizes and several images per launch included."""
    g = torch.Generator().manual_seed(H * 1000 + W)
    img = torch.rand(B, 3, H, W, generator=g)
    Wt = torch.randn(64, 3, 7, 7, generator=g) / 12
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    ref = torch.nn.functional.conv2d(img.double(), Wt.double(), stride=2, padding=3)
    ref = torch.relu(ref * scale.double().view(1, 64, 1, 1) + shift.double().view(1, 64, 1, 1))
    H1, W1 = ref.shape[2], ref.shape[3]
    ref = ref.permute(0, 2, 3, 1).reshape(B * H1 * W1, 64)
    L = _lib.lib()
    s = _lib.cur_stream()
    w7 = torch.zeros(8, 8, 4, 64)
    w7[:7, 1:, :3, :] = Wt.permute(2, 3, 1, 0)
    w7 = w7.reshape(4, 64, 64).contiguous().cuda()
    wmul = 2.0 ** np.floor(np.log2(2048.0 / float(Wt.abs().max())))
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(4, 64, 64, 64)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(w7.data_ptr(), 4, 64, 64, 64, float(wmul), packed.data_ptr(), s))
    sc_d, sh_d, img_d = (scale / wmul).cuda(), shift.cuda(), img.cuda()
    ws_bytes = int(L.imf_image_stem_workspace_bytes(H, W, B))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    Y = torch.zeros(B * H1 * W1, 64, device="cuda")                       # h2 footprint
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(2):
        _lib.check(L.imf_image_stem_h2_fwd(img_d.data_ptr(), H, W, B, packed.data_ptr(), sc_d.data_ptr(), sh_d.data_ptr(), ws.data_ptr(), ws_bytes,
                                           Y.data_ptr(), 128, err.data_ptr(), s))
    out = torch.empty(B * H1 * W1, 64, device="cuda")
    _lib.check(L.imf_h2_unpack(Y.data_ptr(), 128, B * H1 * W1, 64, 64, out.data_ptr(), 64, s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    close(out.cpu().double(), ref, H2_RTOL)


def _p8_pack(x):
    """fp32 [B, H, W, 64] -> the P8 plane buffer (fp16 [B, 16, H + 2, W + 2, 8], zero border) of csrc/image_conv_p8.cu."""
    B, H, W, C = x.shape
    hi = x.half()
    lo = (x - hi.float()).half()
    buf = torch.zeros(B, 16, H + 2, W + 2, 8, dtype=torch.float16)
    buf[:, :8, 1:-1, 1:-1] = hi.reshape(B, H, W, 8, 8).permute(0, 3, 1, 2, 4)
    buf[:, 8:, 1:-1, 1:-1] = lo.reshape(B, H, W, 8, 8).permute(0, 3, 1, 2, 4)
    return buf


def _p8_unpack(buf):
    v = buf[:, :8].float() + buf[:, 8:].float()                       # [B, 8, Hp, Wp, 8]
    assert not v[:, :, 0].any() and not v[:, :, -1].any() and not v[:, :, :, 0].any() and not v[:, :, :, -1].any(), "border written"
    B, _, Hp, Wp, _ = v.shape
    return v[:, :, 1:-1, 1:-1].permute(0, 2, 3, 1, 4).reshape(B, Hp - 2, Wp - 2, 64)


@pytest.mark.parametrize("H,W,B", [(16, 8, 1), (37, 53, 2), (120, 160, 2), (30, 40, 11)])
def test_plane_layout_conv3x3_matches_conv2d(H, W, B):
    """csrc/image_conv_p8.cu: 3x3 / stride 1 / padding 1, 64 -> 64 channels + affine + residual + ReLU on the P8 plane layout against
    torch's conv2d in float64 (model/resnet.py:60-76), partial tiles and more tiles than SMs included; P8 and pixel-major outputs."""
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(B, H, W, 64, generator=g)
    res = torch.randn(B, H, W, 64, generator=g)
    Wt = torch.randn(64, 64, 3, 3, generator=g) / 24
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), Wt.double(), padding=1).permute(0, 2, 3, 1)
    ref_plain = ref * scale.double() + shift.double()
    ref_block = torch.relu(ref_plain + res.double())
    L = _lib.lib()
    s = _lib.cur_stream()
    w9 = Wt.permute(2, 3, 1, 0).reshape(9, 64, 64).contiguous().cuda()
    wmul = 2.0 ** np.floor(np.log2(2048.0 / float(Wt.abs().max())))
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(9, 64, 64, 64)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(w9.data_ptr(), 9, 64, 64, 64, float(wmul), packed.data_ptr(), s))
    sc_d, sh_d = (scale / wmul).cuda(), shift.cuda()
    nbytes = int(L.imf_image_p8_bytes(H, W, B))
    xb, rb = _p8_pack(x).cuda(), _p8_pack(res).cuda()
    assert xb.numel() * 2 == nbytes
    yb = torch.zeros_like(xb)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    # P8 output, no residual, no ReLU
    _lib.check(L.imf_image_conv3x3_p8_fwd(xb.data_ptr(), H, W, B, packed.data_ptr(), sc_d.data_ptr(), sh_d.data_ptr(), None, 0, yb.data_ptr(), 0, 0,
                                          err.data_ptr(), s))
    torch.cuda.synchronize()
    close(_p8_unpack(yb.cpu()).double(), ref_plain, H2_RTOL)
    # pixel-major output with residual + ReLU (what the last block of layer1 does), twice
    ypm = torch.full((B * H * W, 64), float("nan"), device="cuda")          # h2 footprint
    for _ in range(2):
        _lib.check(L.imf_image_conv3x3_p8_fwd(xb.data_ptr(), H, W, B, packed.data_ptr(), sc_d.data_ptr(), sh_d.data_ptr(), rb.data_ptr(), 1,
                                              ypm.data_ptr(), 1, 128, err.data_ptr(), s))
    out = torch.empty(B * H * W, 64, device="cuda")
    _lib.check(L.imf_h2_unpack(ypm.data_ptr(), 128, B * H * W, 64, 64, out.data_ptr(), 64, s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    close(out.cpu().double().reshape(B, H, W, 64), ref_block, H2_RTOL)


def test_plane_layout_maxpool_matches_torch():
    g = torch.Generator().manual_seed(3)
    B, Hin, Win = 2, 45, 62
    x = torch.randn(B, Hin, Win, 64, generator=g)
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    L = _lib.lib()
    s = _lib.cur_stream()
    xd = x.reshape(-1, 64).contiguous().cuda()
    X = torch.zeros(B * Hin * Win, 64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.imf_h2_pack(xd.data_ptr(), 64, B * Hin * Win, 64, 64, X.data_ptr(), 128, err.data_ptr(), s))
    Hout, Wout = ref.shape[1], ref.shape[2]
    yb = torch.zeros(B, 16, Hout + 2, Wout + 2, 8, dtype=torch.float16, device="cuda")
    assert yb.numel() * 2 == int(L.imf_image_p8_bytes(Hout, Wout, B))
    _lib.check(L.imf_image_maxpool_p8(X.data_ptr(), 128, 64, Hin, Win, 3, 2, 1, yb.data_ptr(), B, s))
    torch.cuda.synchronize()
    close(_p8_unpack(yb.cpu()), ref, 1e-6)


def test_module_level_layers_match_standin(frag):
    """imfnet_b200.me layers used one by one (the un-fused route) give the oracle's numbers."""
    import imfnet_b200.me as ME
    coords, ocm, cm = frag
    g = torch.Generator().manual_seed(9)
    X = torch.randn(len(coords), 32, generator=g)
    conv = ME.MinkowskiConvolution(32, 64, kernel_size=3, stride=2, dimension=3).cuda()
    tr = ME.MinkowskiConvolutionTranspose(64, 32, kernel_size=3, stride=2, dimension=3).cuda()
    pw = ME.MinkowskiConvolution(32, 32, kernel_size=1, stride=1, bias=True, dimension=3).cuda()
    x = ME.SparseTensor(X.cuda(), coordinates=torch.from_numpy(coords).cuda())
    y = conv(x)
    z = pw(tr(y))
    ref_y = sparse_ops.conv_forward(X, conv.kernel.detach().cpu(), ocm.table(1, 2, 3, False))
    ref_z = sparse_ops.conv_forward(ref_y, tr.kernel.detach().cpu(), ocm.table(2, 1, 3, True)) @ pw.kernel.detach().cpu() + pw.bias.detach().cpu()
    close(y.F.cpu(), ref_y)
    close(z.F.cpu(), ref_z)
    assert z.coordinate_map_key == x.coordinate_map_key and len(z) == len(x)
    assert ME.cat(x, z).F.shape == (len(coords), 64)


@pytest.mark.parametrize("n,cin,cout,kc_r", [(50000, 64, 64, 64), (50000, 64, 64, 32), (140000, 32, 32, 32), (60000, 32, 64, 64), (50000, 64, 32, 32),
                                             (30000, 64, 128, 64)])
def test_g4_row_mode_residual_matches_simt_conv(n, cin, cout, kc_r):
    """Row mode with a residual: the epilogue takes the residual sub-tiles through the free ring slots (coalesced cp.async, one mbarrier
    per buffer; 140 000 rows at 32 -> 32 = 8 sub-tiles per CTA on 6 buffers, i.e. buffers are re-filled inside a pass), several passes,
    residual chunk width different from the output's, a ragged last tile; 64 -> 128 keeps the direct global-memory reads."""
    coords, _ = synthetic.make_fragment(n, 0.025, 1)
    from imfnet_b200.sparse import CoordinateManager
    cm = CoordinateManager(torch.from_numpy(coords).cuda())
    n = n - 37
    g = torch.Generator(device="cuda").manual_seed(2)
    X = torch.randn(len(coords), cin, device="cuda", generator=g)
    W = torch.randn(27, cin, cout, device="cuda", generator=g) / 40
    R = torch.randn(n, cout, device="cuda", generator=g)
    scale, shift = torch.rand(cout, device="cuda", generator=g) + 0.5, torch.randn(cout, device="cuda", generator=g) * 0.1
    a = run_conv(X, W, cm.table(1, 1, 3, False), n, scale, shift, R, True)
    b = run_conv_g4(X, W, cm.table_t(1, 1, 3, False), n, scale, shift, R, True, split=True, kc_r=kc_r)
    close(b, a, H2_RTOL)


@pytest.fixture(scope="module")
def frag_big():
    """30 000 voxels: more 128-row tiles than SMs at stride 1, i.e. the row-range mode of the persistent kernel."""
    coords, _ = synthetic.make_fragment(30000, 0.025, seed=6)
    ocm = sparse_ops.CoordinateManager(coords)
    ocm.stride(1, 2)
    from imfnet_b200.sparse import CoordinateManager
    cm = CoordinateManager(torch.from_numpy(coords).cuda())
    cm.build_pyramid([2])
    return coords, ocm, cm


def test_parity_grouped_transposed_conv_row_mode(frag_big):
    test_parity_grouped_transposed_conv_matches_oracle(frag_big, 2, 1, 128, 64, True)


@pytest.mark.parametrize("t_in,t_out,cin,cout", [(2, 1, 128, 64), (4, 2, 256, 64), (8, 4, 256, 128)])
@pytest.mark.parametrize("split", [False, True])
def test_parity_grouped_transposed_conv_matches_oracle(frag, t_in, t_out, cin, cout, split):
    """imf_parity_perm + imf_kernel_map_t_batch(perm) + imf_sparse_conv_g4_fwd_perm: the transposed convolution computed in
    parity-grouped row order gives the oracle's result in the original row order, and its tiles need few offsets."""
    coords, ocm, cm = frag
    L = _lib.lib()
    fine, coarse = cm.level(t_out), cm.level(t_in)
    n = fine.n
    ld_n = (n + 127) // 128 * 128
    s = _lib.cur_stream()
    perm = torch.full((ld_n,), -7, dtype=torch.int32, device="cuda")
    ws_b = int(L.imf_parity_perm_workspace_bytes(n))
    ws_p = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_parity_perm(fine.coords.data_ptr(), None, n, t_out, perm.data_ptr(), ws_p.data_ptr(), ws_b, s))
    p = perm[:n].cpu().numpy()
    assert sorted(p.tolist()) == list(range(n))                                   # a permutation
    c = fine.coords.cpu().numpy()
    cls = ((c[:, 1] // t_out) & 1) | (((c[:, 2] // t_out) & 1) << 1) | (((c[:, 3] // t_out) & 1) << 2)
    assert np.array_equal(p, np.argsort(cls, kind="stable"))                      # grouped by class, stable inside
    nbr_t = torch.empty((27, ld_n), dtype=torch.int32, device="cuda")
    mask = torch.empty(ld_n // 128 + 1, dtype=torch.int32, device="cuda")
    job = (_lib.KmapJob * 1)(_lib.KmapJob(fine.coords.data_ptr(), None, coarse.table.data_ptr(), nbr_t.data_ptr(), mask.data_ptr(),
                                          perm.data_ptr(), -t_out, None, None))
    _lib.check(L.imf_kernel_map_t_batch(job, 1, n, coarse.capacity, 3, ld_n, s))
    onbr = ocm.table(t_in, t_out, 3, True)
    assert np.array_equal(nbr_t.cpu().numpy()[:, :n], onbr[p].T)
    offsets_per_tile = [bin(int(m) & 0x7FFFFFF).count("1") for m in mask.cpu().numpy()[: (n + 127) // 128]]
    assert np.mean(offsets_per_tile) < 14                                          # 27 without the grouping
    g = torch.Generator().manual_seed(cin + cout)
    X = torch.randn(coarse.n, cin, generator=g)
    W = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = torch.relu(sparse_ops.conv_forward(X, W, onbr) * scale + shift)
    kc_in, kc_out = 64, 64
    wmul = 2.0 ** np.floor(np.log2(2048.0 / float(W.abs().max())))
    Wd = W.cuda()
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(27, cin, cout, kc_in)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(Wd.data_ptr(), 27, cin, cout, kc_in, wmul, packed.data_ptr(), s))
    Xh = h2_pack(X.cuda(), kc_in)
    Yh = torch.full((n, 2 * cout), float("nan"), dtype=torch.float16, device="cuda")
    ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(cout)) if split else 0
    ws = torch.zeros(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    sc, sh = (scale / wmul).cuda().contiguous(), shift.cuda()
    _lib.check(L.imf_sparse_conv_g4_fwd_perm(Xh.data_ptr(), Xh.stride(0), kc_in, packed.data_ptr(), nbr_t.data_ptr(), ld_n, mask.data_ptr(),
                                             None, n, 27, cin, cout, sc.data_ptr(), sh.data_ptr(), None, 0, 0, 1, Yh.data_ptr(),
                                             Yh.stride(0), n, kc_out, perm.data_ptr(), ws.data_ptr() if split else None, ws_bytes,
                                             err.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    close(h2_unpack(Yh, cout, kc_out).cpu(), ref, H2_RTOL)
