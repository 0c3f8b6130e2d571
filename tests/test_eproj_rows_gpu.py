"""The small-edge-set variant of the edge-attribute projection (csrc/dense.cu::ctan_eproj_fwd_rows) against the
tiled contraction it stands in for (ctan_eproj_fwd) and an fp64 evaluation of
edge_attr = [msg[e_id] | cos(w * rel + b)] -> lin_edge (models/gnn_layers.py:97, TGB/modules/time_enc.py:23-24)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E,Fe,T,D,use_idx,e_true", [
    (1106, 172, 16, 128, True, None),      # the wiki training batch
    (5, 4, 4, 32, True, None),             # fewer rows than one group
    (3000, 172, 16, 128, True, 1111),      # device-sized row count below the bound; rows past it untouched
    (700, 6, 2, 64, True, None),           # message width not a multiple of 4: scalar staging
    (257, 100, 100, 256, False, None),     # identity row map (TGB evaluation layout), D = 256
    (64, 8, 8, 128, True, 0),              # empty on the device
])
def test_eproj_rows_matches_tiled_and_fp64(E, Fe, T, D, use_idx, e_true):
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(E + Fe + D)
    n_rows = 5000
    msg = torch.randn(n_rows if use_idx else E, Fe, generator=g).to(dev)
    eidx = torch.randint(0, n_rows, (E,), generator=g).to(dev) if use_idx else None
    rel = (torch.rand(E, generator=g) * 50.0).to(dev)
    tw, tb = torch.randn(T, generator=g).to(dev), torch.randn(T, generator=g).to(dev)
    We = (torch.randn(D, Fe + T, generator=g) / (Fe + T) ** 0.5).to(dev)
    Ed = None if e_true is None else torch.tensor([e_true], dtype=torch.int32, device=dev)
    n = E if e_true is None else e_true
    a = torch.full((E, D), 3.0, device=dev)
    b = torch.full((E, D), 3.0, device=dev)
    args = (ptr(msg), Fe, ptr(eidx) if use_idx else None, ptr(rel), ptr(tw), ptr(tb), T, ptr(We), E,
            ptr(Ed) if Ed is not None else None, D)
    call("ctan_eproj_fwd", *args, ptr(a))
    call("ctan_eproj_fwd_rows", *args, ptr(b))
    torch.cuda.synchronize()
    assert bool((b[n:] == 3.0).all())
    if n == 0:
        return
    rows = msg[eidx[:n]] if use_idx else msg[:n]
    tenc = torch.cos(rel[:n, None].double() * tw.double() + tb.double())
    ref = torch.cat([rows.double(), tenc], 1) @ We.double().t()
    s = float(ref.abs().max())
    assert float((a[:n].double() - ref).abs().max()) < 1e-5 * s
    assert float((b[:n].double() - ref).abs().max()) < 1e-5 * s
    assert float((a[:n] - b[:n]).abs().max()) < 2e-6 * s


def test_eproj_rows_rejects_unsupported_shapes():
    from ctan_b200._lib import CtanLibraryError, call, ptr
    dev = torch.device("cuda:0")
    z = torch.zeros(64, 74, device=dev)
    with pytest.raises(CtanLibraryError):                      # D = 74: the caller keeps the tiled contraction
        call("ctan_eproj_fwd_rows", ptr(z), 4, None, ptr(z), ptr(z), ptr(z), 4, ptr(z), 8, None, 74, ptr(z))
    with pytest.raises(CtanLibraryError):                      # Fe + T = 7
        call("ctan_eproj_fwd_rows", ptr(z), 4, None, ptr(z), ptr(z), ptr(z), 3, ptr(z), 8, None, 64, ptr(z))


@pytest.mark.parametrize("E,Fe,T,D,e_true", [(1106, 172, 16, 128, None), (3000, 8, 100, 64, 777), (40, 4, 4, 32, None),
                                              (64, 8, 8, 128, 0)])
def test_tenc_bwd_fused_matches_two_kernel_path_and_fp64(E, Fe, T, D, e_true):
    """Time-encoder gradients straight from dEP (ctan_tenc_bwd_fused) against ctan_eproj_bwd(.., dTenc) +
    ctan_tenc_bwd and against fp64 autograd of cos(w * rel + b) -> lin_edge (TGB/modules/time_enc.py:23-24)."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(E + T + D)
    dEP = torch.randn(E, D, generator=g).to(dev)
    We = (torch.randn(D, Fe + T, generator=g) / (Fe + T) ** 0.5).to(dev)
    rel = (torch.rand(E, generator=g) * 20.0).to(dev)
    tw, tb = torch.randn(T, generator=g).to(dev), torch.randn(T, generator=g).to(dev)
    msg = torch.randn(E, Fe, generator=g).to(dev)
    Ed = None if e_true is None else torch.tensor([e_true], dtype=torch.int32, device=dev)
    n = E if e_true is None else e_true
    a_w, a_b = torch.zeros(T, device=dev), torch.zeros(T, device=dev)
    b_w, b_b = torch.zeros(T, device=dev), torch.zeros(T, device=dev)
    dT = torch.zeros(E, T, device=dev)
    edp = ptr(Ed) if Ed is not None else None
    call("ctan_eproj_bwd", ptr(dEP), ptr(msg), Fe, None, ptr(rel), ptr(tw), ptr(tb), T, ptr(We), E, edp, D, None, ptr(dT), 0)
    call("ctan_tenc_bwd", ptr(dT), ptr(rel), ptr(tw), ptr(tb), E, edp, T, ptr(a_w), ptr(a_b))
    call("ctan_tenc_bwd_fused", ptr(dEP), ptr(We), Fe, T, D, ptr(rel), ptr(tw), ptr(tb), E, edp, ptr(b_w), ptr(b_b))
    torch.cuda.synchronize()
    if n == 0:
        assert float(b_w.abs().max()) == 0.0 and float(b_b.abs().max()) == 0.0
        return
    w64, b64 = tw.double().requires_grad_(), tb.double().requires_grad_()
    tenc = torch.cos(rel[:n, None].double() * w64 + b64)
    (tenc @ We[:, Fe:].double().t() * dEP[:n].double()).sum().backward()
    for got, ref in ((a_w, w64.grad), (b_w, w64.grad), (a_b, b64.grad), (b_b, b64.grad)):
        assert float((got.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-6, (got, ref)


@pytest.mark.parametrize("N,D,n_true", [(640, 128, None), (3600, 128, 611), (200, 32, None)])
def test_qkva_wgrad_matches_bwd_plus_antisym(N, D, n_true):
    """ctan_qkva_wgrad (dA - dA^T accumulated straight into dW) against ctan_qkva_bwd + ctan_antisym_bwd and fp64."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N + D)
    DQ = torch.randn(N, 4 * D, generator=g).to(dev)
    X = torch.randn(N, D, generator=g).to(dev)
    Nd = None if n_true is None else torch.tensor([n_true], dtype=torch.int32, device=dev)
    n = N if n_true is None else n_true
    ndp = ptr(Nd) if Nd is not None else None
    A = [torch.zeros(D, D, device=dev) for _ in range(4)] + [torch.zeros(D, device=dev) for _ in range(4)]
    B = [torch.zeros(D, D, device=dev) for _ in range(4)] + [torch.zeros(D, device=dev) for _ in range(4)]
    dW_a = torch.zeros(D, D, device=dev)
    z = torch.zeros(D, D, device=dev)
    call("ctan_qkva_bwd", ptr(DQ), ptr(X), None, N, ndp, D, ptr(z), ptr(z), ptr(z), ptr(z), None, *(ptr(t) for t in A),
         min(N, 900))
    call("ctan_antisym_bwd", ptr(A[3]), D, ptr(dW_a))
    call("ctan_qkva_wgrad", ptr(DQ), ptr(X), N, ndp, D, *(ptr(t) for t in B), min(N, 900))
    torch.cuda.synchronize()
    ref = DQ[:n].double().t() @ X[:n].double()                    # [4D, D]
    refs = [ref[i * D:(i + 1) * D] for i in range(3)] + [ref[3 * D:] - ref[3 * D:].t()]
    for i in range(4):
        got_a = dW_a if i == 3 else A[i]
        s = float(refs[i].abs().max())
        assert float((got_a.double() - refs[i]).abs().max()) <= 1e-5 * s
        assert float((B[i].double() - refs[i]).abs().max()) <= 1e-5 * s
        bref = DQ[:n, i * D:(i + 1) * D].double().sum(0)
        assert float((B[4 + i].double() - bref).abs().max()) <= 1e-5 * float(bref.abs().max())
        assert float((A[4 + i] - B[4 + i]).abs().max()) <= 1e-5 * float(bref.abs().max())


@pytest.mark.parametrize("P,n_pos,D,Hd,N", [(400, 200, 128, 64, 640), (401, 200, 128, 64, 300), (30, 10, 32, 32, 50),
                                            (66, 33, 128, 128, 100), (10, 5, 64, 256, 20)])
def test_readout_link_fused_matches_three_kernel_path_and_autograd(P, n_pos, D, Hd, N):
    """ctan_readout_link_fused against ctan_readout_fwd + ctan_head_loss + ctan_readout_bwd and against fp64 autograd of
    LinkPredictor (models/predictors.py:16-19) + BCEWithLogitsLoss on positives / negatives (train_link.py:269-270)."""
    from ctan_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(P + D + Hd)
    rnd = lambda *s: torch.randn(*s, generator=g)
    Z = rnd(N, D).to(dev)
    n_glob = 3 * N
    perm = torch.randperm(n_glob, generator=g)[:N]
    amap = torch.full((n_glob,), -1, dtype=torch.long)
    amap[perm] = torch.arange(N)
    src = perm[torch.randint(0, N, (P,), generator=g)].to(dev)
    dst = perm[torch.randint(0, N, (P,), generator=g)].to(dev)
    amap = amap.to(dev)
    Wsrc, Wdst = (rnd(Hd, D) / D ** 0.5).to(dev), (rnd(Hd, D) / D ** 0.5).to(dev)
    bsrc, bdst = rnd(Hd).to(dev), rnd(Hd).to(dev)
    Wfin, bfin = (rnd(1, Hd) / Hd ** 0.5).to(dev), rnd(1).to(dev)
    ctr = torch.tensor([0, 0, 0, 1], dtype=torch.long, device=dev)

    def bufs():
        return dict(H=torch.zeros(P, Hd, device=dev), OUT=torch.zeros(P, device=dev), dOUT=torch.zeros(P, device=dev),
                    dHid=torch.zeros(P, Hd, device=dev), dZ=torch.zeros(N, D, device=dev), loss=torch.zeros(4, device=dev))
    a, b = bufs(), bufs()
    call("ctan_readout_fwd", ptr(Z), D, ptr(src), ptr(dst), ptr(amap), P, None, Hd, 1, ptr(Wsrc), ptr(bsrc), ptr(Wdst),
         ptr(bdst), ptr(Wfin), ptr(bfin), ptr(a["H"]), None, 1)
    call("ctan_head_loss", ptr(a["H"]), P, Hd, ptr(Wfin), ptr(bfin), n_pos, ptr(a["OUT"]), ptr(a["dOUT"]), ptr(a["dHid"]),
         ptr(a["loss"]), 4, ptr(ctr))
    call("ctan_readout_bwd", ptr(a["dOUT"]), ptr(a["H"]), ptr(Z), D, ptr(src), ptr(dst), ptr(amap), P, None, Hd, 1,
         ptr(Wsrc), ptr(Wdst), ptr(Wfin), ptr(a["dHid"]), ptr(a["dZ"]), None, None, None, None, None, None, 1)
    call("ctan_readout_link_fused", ptr(Z), D, ptr(src), ptr(dst), ptr(amap), P, n_pos, Hd, ptr(Wsrc), ptr(bsrc),
         ptr(Wdst), ptr(bdst), ptr(Wfin), ptr(bfin), ptr(b["H"]), ptr(b["OUT"]), ptr(b["dOUT"]), ptr(b["dHid"]),
         ptr(b["dZ"]), ptr(b["loss"]), 4, ptr(ctr))
    # no-grad form: scores + loss only
    c = bufs()
    call("ctan_readout_link_fused", ptr(Z), D, ptr(src), ptr(dst), ptr(amap), P, n_pos, Hd, ptr(Wsrc), ptr(bsrc),
         ptr(Wdst), ptr(bdst), ptr(Wfin), ptr(bfin), ptr(c["H"]), ptr(c["OUT"]), None, None, None, ptr(c["loss"]), 4,
         ptr(ctr))
    torch.cuda.synchronize()
    z64 = Z.double().requires_grad_()
    rs, rd = amap[src], amap[dst]
    h = z64[rs] @ Wsrc.double().t() + bsrc.double() + z64[rd] @ Wdst.double().t() + bdst.double()
    out = (h.relu() @ Wfin.double().t() + bfin.double()).squeeze(-1)
    bce = torch.nn.BCEWithLogitsLoss()
    loss = bce(out[:n_pos], torch.ones(n_pos, dtype=torch.float64, device=dev)) + \
        bce(out[n_pos:], torch.zeros(P - n_pos, dtype=torch.float64, device=dev))
    loss.backward()
    for d in (a, b):
        assert float((d["OUT"].double() - out.detach()).abs().max()) <= 2e-5 * float(out.abs().max()) + 1e-6
        assert abs(float(d["loss"][0]) - float(loss)) <= 1e-5 * abs(float(loss))
        assert float((d["dZ"].double() - z64.grad).abs().max()) <= 2e-5 * float(z64.grad.abs().max()) + 1e-9
    assert float((a["H"] - b["H"]).abs().max()) <= 1e-5 * float(a["H"].abs().max())
    assert float((a["dHid"] - b["dHid"]).abs().max()) <= 1e-5 * float(a["dHid"].abs().max()) + 1e-9
    assert float((a["dOUT"] - b["dOUT"]).abs().max()) <= 1e-6
    assert float((c["OUT"] - b["OUT"]).abs().max()) == 0.0 and abs(float(c["loss"][0] - b["loss"][0])) <= 1e-6
    assert float(c["dZ"].abs().max()) == 0.0
