"""tcgen05/TMA projection (csrc/tc_gemm.cu, 3xTF32) against the CUDA-core fp32 contraction and a float64
reference: it must stay inside the fp32 path's 1e-4 bar (north_star), far tighter than plain TF32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D", [(700, 128), (5000, 256), (129, 64), (128, 32)])
def test_qkva_tc_matches_fp32(N, D):
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N + D)
    X = torch.randn(N, D, generator=g).to(dev)
    Ws = [(torch.randn(D, D, generator=g) / D ** 0.5).to(dev) for _ in range(4)]     # Wq, Wk, Wv, W
    bs = [torch.randn(D, generator=g).to(dev) for _ in range(4)]
    gamma = 0.1
    hi = torch.empty(4 * D, D, device=dev)
    lo = torch.empty(4 * D, D, device=dev)
    A = torch.empty(D, D, device=dev)
    call("ctan_wcat_prep", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), gamma, D, ptr(hi), ptr(lo), ptr(A), None, None)
    A_ref = Ws[3] - Ws[3].t() - gamma * torch.eye(D, device=dev)
    assert torch.equal(A, A_ref)
    assert torch.equal(hi + lo, torch.cat([Ws[0], Ws[1], Ws[2], A_ref]))                # exact split
    out_tc = torch.full((N, 4 * D), float("nan"), device=dev)
    call("ctan_qkva_fwd_tc", ptr(X), N, None, D, ptr(hi), ptr(lo), ptr(bs[0]), ptr(bs[1]), ptr(bs[2]), ptr(bs[3]),
         ptr(out_tc), 0)
    out_fp = torch.empty((N, 4 * D), device=dev)
    call("ctan_qkva_fwd", ptr(X), N, None, D, ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(A), ptr(bs[0]), ptr(bs[1]),
         ptr(bs[2]), ptr(bs[3]), ptr(out_fp))
    torch.cuda.synchronize()
    Wcat = torch.cat([Ws[0], Ws[1], Ws[2], A_ref]).double()
    ref = X.double() @ Wcat.t() + torch.cat(bs).double()
    scale = float(ref.abs().max())
    e_tc = float((out_tc.double() - ref).abs().max()) / scale
    e_fp = float((out_fp.double() - ref).abs().max()) / scale
    assert not torch.isnan(out_tc).any()
    assert e_fp < 2e-6, e_fp
    assert e_tc < 5e-6, (e_tc, e_fp)          # 3xTF32: ~fp32 accuracy (plain TF32 would be ~5e-4)
    out_1x = torch.full((N, 4 * D), float("nan"), device=dev)
    call("ctan_qkva_fwd_tc", ptr(X), N, None, D, ptr(hi), ptr(lo), ptr(bs[0]), ptr(bs[1]), ptr(bs[2]), ptr(bs[3]),
         ptr(out_1x), 1)
    torch.cuda.synchronize()
    e_1x = float((out_1x.double() - ref).abs().max()) / scale
    assert 1e-5 < e_1x < 5e-3, e_1x           # single-pass TF32: the 1e-2 path


def test_qkva_tc_device_sized_rows():
    """N read on the device (engine mode): rows past the true N are neither computed nor written."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    D, cap, n_true = 128, 1024, 300
    X = torch.randn(cap, D, device=dev)
    W = [torch.randn(D, D, device=dev) / 11 for _ in range(4)]
    hi, lo = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev)
    call("ctan_wcat_prep", ptr(W[0]), ptr(W[1]), ptr(W[2]), ptr(W[3]), 0.1, D, ptr(hi), ptr(lo), None, None, None)
    out = torch.full((cap, 4 * D), 7.0, device=dev)
    n_dev = torch.tensor([n_true], dtype=torch.int32, device=dev)
    call("ctan_qkva_fwd_tc", ptr(X), cap, ptr(n_dev), D, ptr(hi), ptr(lo), None, None, None, None, ptr(out), 0)
    torch.cuda.synchronize()
    ref = X[:n_true].double() @ (hi + lo).double().t()
    assert float((out[:n_true].double() - ref).abs().max()) < 5e-5
    assert bool((out[384:] == 7.0).all())           # whole tiles past n_true untouched
    assert bool((out[n_true:384] == 7.0).all())     # rows past n_true inside the last tile untouched


@pytest.mark.parametrize("N,D", [(700, 128), (1300, 256)])
def test_dx_tc_matches_fp32(N, D):
    """dX = G + DQKVA Wcat on tcgen05 vs the CUDA-core split-K contraction and float64."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N)
    DQ = torch.randn(N, 4 * D, generator=g).to(dev)
    G = torch.randn(N, D, generator=g).to(dev)
    Ws = [(torch.randn(D, D, generator=g) / D ** 0.5).to(dev) for _ in range(4)]
    hi, lo = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev)
    hiT, loT = torch.empty(D, 4 * D, device=dev), torch.empty(D, 4 * D, device=dev)
    A = torch.empty(D, D, device=dev)
    call("ctan_wcat_prep", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), 0.1, D, ptr(hi), ptr(lo), ptr(A), ptr(hiT),
         ptr(loT))
    assert torch.equal(hiT, hi.t().contiguous()) and torch.equal(loT, lo.t().contiguous())
    out_tc = torch.full((N, D), float("nan"), device=dev)
    call("ctan_dx_tc", ptr(DQ), ptr(G), N, None, D, ptr(hiT), ptr(loT), ptr(out_tc), 0, 1)
    out_sp = torch.zeros((N, D), device=dev)                      # split-K variant accumulates into zeros
    call("ctan_dx_tc", ptr(DQ), ptr(G), N, None, D, ptr(hiT), ptr(loT), ptr(out_sp), 0, 4)
    out_fp = torch.zeros((N, D), device=dev)
    call("ctan_qkva_bwd", ptr(DQ), ptr(DQ), ptr(G), N, None, D, ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(A), ptr(out_fp),
         None, None, None, None, None, None, None, None, 0)
    torch.cuda.synchronize()
    ref = G.double() + DQ.double() @ (hi + lo).double()
    scale = float(ref.abs().max())
    assert float((out_fp.double() - ref).abs().max()) / scale < 2e-6
    assert float((out_tc.double() - ref).abs().max()) / scale < 1e-5      # K = 4D up to 1024 terms
    assert float((out_sp.double() - ref).abs().max()) / scale < 1e-5


def test_qkva_tc_persistent_row_loop():
    """Device-sized N far below a large bound: the launch is capped and each CTA walks several row tiles
    (smem ring phases and the TMEM accumulator hand-over run across tiles)."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    D, cap, n_true = 128, 60000, 41000
    X = torch.randn(cap, D, device=dev)
    W = [torch.randn(D, D, device=dev) / 11 for _ in range(4)]
    b = [torch.randn(D, device=dev) for _ in range(4)]
    hi, lo = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev)
    call("ctan_wcat_prep", ptr(W[0]), ptr(W[1]), ptr(W[2]), ptr(W[3]), 0.1, D, ptr(hi), ptr(lo), None, None, None)
    out = torch.full((cap, 4 * D), 7.0, device=dev)
    n_dev = torch.tensor([n_true], dtype=torch.int32, device=dev)
    call("ctan_qkva_fwd_tc", ptr(X), cap, ptr(n_dev), D, ptr(hi), ptr(lo), ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]),
         ptr(out), 0)
    torch.cuda.synchronize()
    ref = X[:n_true].double() @ (hi + lo).double().t() + torch.cat(b).double()
    err = float((out[:n_true].double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 5e-6, err
    assert bool((out[n_true:] == 7.0).all())


@pytest.mark.parametrize("M,K,Nout", [(300, 64, 128), (1000, 256, 512), (4200, 320, 256)])
def test_bf16_mode_matches_fp64_at_1e2(M, K, Nout):
    """mode 2 (the north star's 1e-2 path): activations converted to BF16 in the kernel, BF16 weight panel, fp32
    accumulation in TMEM; bias and residual epilogue as in the other modes; device-sized M."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn(M + 37, K, generator=g).to(dev)
    W = (torch.randn(Nout, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(Nout, generator=g).to(dev)
    res = torch.randn(M + 37, Nout, generator=g).to(dev)
    Wb = torch.zeros((Nout, K), dtype=torch.bfloat16, device=dev)
    call("ctan_pack_bf16", ptr(W), Nout, K, K, K, ptr(Wb))
    assert torch.equal(Wb, W.to(torch.bfloat16))
    C = torch.full((M + 37, Nout), 7.0, device=dev)
    Md = torch.tensor([M], dtype=torch.int32, device=dev)
    call("ctan_gemm_tc", ptr(A), M + 37, ptr(Md), K, ptr(Wb), None, Nout, ptr(b), None, None, None, Nout, ptr(res), Nout,
         ptr(C), 2)
    ref = A[:M].double() @ W.double().t() + b.double() + res[:M].double()
    err = float((C[:M].double() - ref).abs().max() / ref.abs().max())
    assert 1e-5 < err < 1e-2, err
    assert bool((C[M:] == 7.0).all())                       # rows past the true M untouched
    exact = A[:M].to(torch.bfloat16).double() @ Wb.double().t() + b.double() + res[:M].double()
    assert float((C[:M].double() - exact).abs().max() / ref.abs().max()) < 1e-5     # = bf16 inputs, fp32 accumulate


@pytest.mark.parametrize("N,D,Fn", [(640, 128, 1), (300, 128, 0), (1000, 256, 5), (100, 96, 3), (100, 32, 0)])
def test_folded_enc_projection_matches_fp64(N, D, Fn):
    """enc_x folded into the projection (csrc/tc_gemm.cu): the per-step panel kernel ctan_fuse_enc_prep_split against
    the once-per-run pair (ctan_fuse_enc_prep + ctan_pack_split) and fp64, then [X0 | QKVA] = zext Wce^T + bce on
    tcgen05 against the two chained fp64 maps X0 = z Wenc^T + benc, QKVA = X0 Wcat^T + bcat
    (models/gnn_layers.py:96 + PyG TransformerConv / AntiSymmetricConv projections, App. A.5)."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N + D + Fn)
    rnd = lambda *s: torch.randn(*s, generator=g)
    Ws = [(rnd(D, D) / D ** 0.5).to(dev) for _ in range(4)]                     # Wq, Wk, Wv, W
    bs = [rnd(D).to(dev) for _ in range(4)]
    Wenc, benc = (rnd(D, D + Fn) / D ** 0.5).to(dev), rnd(D).to(dev)
    gamma = 0.1
    Kx = ((D + Fn + 31) // 32) * 32
    A = (Ws[3] - Ws[3].t() - gamma * torch.eye(D, device=dev)).contiguous()
    Wce, bce = torch.empty(5 * D, Kx, device=dev), torch.empty(5 * D, device=dev)
    hi1, lo1 = torch.empty_like(Wce), torch.empty_like(Wce)
    call("ctan_fuse_enc_prep", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(A), ptr(bs[0]), ptr(bs[1]), ptr(bs[2]), ptr(bs[3]),
         ptr(Wenc), ptr(benc), D, Fn, Kx, ptr(Wce), ptr(bce))
    call("ctan_pack_split", ptr(Wce), 5 * D, Kx, Kx, ptr(hi1), ptr(lo1))
    hi2, lo2 = torch.full_like(Wce, float("nan")), torch.full_like(Wce, float("nan"))
    bce2 = torch.full_like(bce, float("nan"))
    call("ctan_fuse_enc_prep_split", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), gamma, ptr(bs[0]), ptr(bs[1]),
         ptr(bs[2]), ptr(bs[3]), ptr(Wenc), ptr(benc), D, Fn, Kx, ptr(hi2), ptr(lo2), ptr(bce2))
    torch.cuda.synchronize()
    Wcat = torch.cat([Ws[0], Ws[1], Ws[2], A]).double()
    Wce_ref = torch.zeros(5 * D, Kx, dtype=torch.float64, device=dev)
    Wce_ref[:D, :D + Fn] = Wenc.double()
    Wce_ref[D:, :D + Fn] = Wcat @ Wenc.double()
    bce_ref = torch.cat([benc.double(), Wcat @ benc.double() + torch.cat(bs).double()])
    s = float(Wce_ref.abs().max())
    assert float(((hi1 + lo1).double() - Wce_ref).abs().max()) < 1e-6 * s
    assert not torch.isnan(hi2).any() and not torch.isnan(lo2).any() and not torch.isnan(bce2).any()
    assert float(((hi2 + lo2).double() - Wce_ref).abs().max()) < 2e-6 * s          # fp32 accumulation
    assert torch.equal(hi2[:D] + lo2[:D], Wce[:D])                                 # the enc_x block is a copy
    assert float((bce2.double() - bce_ref).abs().max()) < 2e-6 * float(bce_ref.abs().max())
    if D % 128:
        return                                       # (the panel kernel is general in D % 32; the contraction needs 128-wide tiles)
    # the contraction itself, device-sized row count, rows beyond N untouched
    mem, xfeat = rnd(2000, D).to(dev), (rnd(2000, Fn).to(dev) if Fn else None)
    n_id = torch.randperm(2000, generator=g)[:N].sort().values.to(dev)
    Nd = torch.tensor([N], dtype=torch.int32, device=dev)
    cap = N + 200
    zext = torch.zeros(cap + 128, Kx, device=dev)
    call("ctan_enc_prep_ext", ptr(mem), D, ptr(xfeat) if Fn else None, Fn, ptr(n_id), cap, ptr(Nd), Kx, ptr(zext))
    X0 = torch.full((cap, D), 7.0, device=dev)
    Q = torch.full((cap, 4 * D), 7.0, device=dev)
    call("ctan_enc_qkva_tc", ptr(zext), cap, ptr(Nd), Kx, ptr(hi2), ptr(lo2), D, ptr(bce2), ptr(X0), ptr(Q), 0 | (N << 8))
    torch.cuda.synchronize()
    z = torch.cat([mem[n_id], xfeat[n_id]], 1).double() if Fn else mem[n_id].double()
    X0_ref = z @ Wenc.double().t() + benc.double()
    Q_ref = X0_ref @ Wcat.t() + torch.cat(bs).double()
    assert float((X0[:N].double() - X0_ref).abs().max()) < 5e-6 * float(X0_ref.abs().max())
    assert float((Q[:N].double() - Q_ref).abs().max()) < 1e-5 * float(Q_ref.abs().max())
    assert bool((X0[N:] == 7.0).all()) and bool((Q[N:] == 7.0).all())


@pytest.mark.timeout(300)
@pytest.mark.parametrize("N,D,n_true", [(700, 128, None), (5000, 256, None), (129, 128, None), (300, 128, None),
                                        (40000, 256, None), (3000, 128, 1300), (128, 128, None)])
def test_tc_pair_multicast_is_bit_identical(N, D, n_true):
    """The two-CTA-cluster form of the tcgen05 kernel (weight tiles fetched once per pair and TMA-multicast, `mode` bit 6)
    issues exactly the same MMAs per tile as the single-CTA form: outputs must be BIT-identical -- odd row-tile counts
    (the second CTA of the last pair takes a dummy tile), device-sized row counts and the folded enc_x + projection form
    (column redirect) included."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N + D)
    X = torch.randn(N, D, generator=g).to(dev)
    Ws = [(torch.randn(D, D, generator=g) / D ** 0.5).to(dev) for _ in range(4)]
    bs = [torch.randn(D, generator=g).to(dev) for _ in range(4)]
    hi, lo, A = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev), torch.empty(D, D, device=dev)
    call("ctan_wcat_prep", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), 0.1, D, ptr(hi), ptr(lo), ptr(A), None, None)
    Nd = None if n_true is None else torch.tensor([n_true], dtype=torch.int32, device=dev)
    n = N if n_true is None else n_true
    outs = []
    for mode in (0, 0x40):
        out = torch.full((N, 4 * D), 5.0, device=dev)
        call("ctan_qkva_fwd_tc", ptr(X), N, ptr(Nd) if Nd is not None else None, D, ptr(hi), ptr(lo), ptr(bs[0]), ptr(bs[1]),
             ptr(bs[2]), ptr(bs[3]), ptr(out), mode | ((n << 8) if n_true else 0))
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert bool((outs[1][n:] == 5.0).all())
    ref = X[:n].double() @ torch.cat([Ws[0], Ws[1], Ws[2], A]).double().t() + torch.cat(bs).double()
    assert float((outs[1][:n].double() - ref).abs().max()) < 5e-6 * float(ref.abs().max())
    # generic contraction with a residual + the folded form with the column redirect
    R = torch.randn(N, 4 * D, generator=g).to(dev)
    o2 = []
    for mode in (0, 0x40):
        out = torch.full((N, 4 * D), 5.0, device=dev)
        call("ctan_gemm_tc", ptr(X), N, ptr(Nd) if Nd is not None else None, D, ptr(hi), ptr(lo), 4 * D, ptr(bs[0]), ptr(bs[1]),
             ptr(bs[2]), ptr(bs[3]), D, ptr(R), 4 * D, ptr(out), mode)
        o2.append(out)
    torch.cuda.synchronize()
    assert torch.equal(o2[0], o2[1])
    if D % 128 == 0:
        Fn = 3
        Kx = ((D + Fn + 31) // 32) * 32
        Wenc, benc = (torch.randn(D, D + Fn, generator=g) / D ** 0.5).to(dev), torch.randn(D, generator=g).to(dev)
        ph, pl, bce = torch.empty(5 * D, Kx, device=dev), torch.empty(5 * D, Kx, device=dev), torch.empty(5 * D, device=dev)
        call("ctan_fuse_enc_prep_split", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), 0.1, ptr(bs[0]), ptr(bs[1]), ptr(bs[2]),
             ptr(bs[3]), ptr(Wenc), ptr(benc), D, Fn, Kx, ptr(ph), ptr(pl), ptr(bce))
        zext = torch.zeros(N + 128, Kx, device=dev)
        zext[:N, :D + Fn] = torch.randn(N, D + Fn, generator=g).to(dev)
        o3 = []
        for mode in (0, 0x40):
            X0, Q = torch.full((N, D), 5.0, device=dev), torch.full((N, 4 * D), 5.0, device=dev)
            call("ctan_enc_qkva_tc", ptr(zext), N, ptr(Nd) if Nd is not None else None, Kx, ptr(ph), ptr(pl), D, ptr(bce), ptr(X0),
                 ptr(Q), mode)
            o3.append((X0, Q))
        torch.cuda.synchronize()
        assert torch.equal(o3[0][0], o3[1][0]) and torch.equal(o3[0][1], o3[1][1])
