"""ctan_gemm_mma (mma.sync m16n8k8 TF32, 3xTF32 compensation) against a float64 torch reference of the same
contraction: fp32-level agreement (1e-5 of the output scale; plain fp32 FMA gives ~1e-6), every operand mode the
engines use (segmented W^T weights + bias, N-contiguous weights + residual + split-K, gathered rows + node-feature
tail, device-side row count, ragged M)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _call(A, lda, a_idx, M, M_dev, K, Bs, ldb, seg, b_nn, Nout, biases, bseg, res, xf, Fn, xw, ldxw, C, n_split, mode=0):
    from ctan_b200._lib import call, ptr
    Bp = [ptr(b) for b in Bs] + [None] * (4 - len(Bs))
    bp = [ptr(b) if b is not None else None for b in biases] + [None] * (4 - len(biases))
    call("ctan_gemm_mma", ptr(A), lda, ptr(a_idx) if a_idx is not None else None, M, ptr(M_dev) if M_dev is not None else None,
         K, *Bp, ldb, seg, b_nn, Nout, *bp, bseg, ptr(res) if res is not None else None, res.size(1) if res is not None else 0,
         ptr(xf) if xf is not None else None, Fn, xw, ldxw, ptr(C), C.size(1), n_split, mode)


def _err(c, ref):
    return float((c.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M", [1, 63, 700, 1029])
def test_projection_form_four_segments_bias(M):
    g = torch.Generator().manual_seed(M)
    D = 128
    X = torch.randn(M, D, generator=g).to(DEV)
    W = [torch.randn(D, D, generator=g).to(DEV) * 0.1 for _ in range(4)]
    b = [torch.randn(D, generator=g).to(DEV) for _ in range(4)]
    C = torch.full((M, 4 * D), float("nan"), device=DEV)
    _call(X, D, None, M, None, D, W, D, D, 0, 4 * D, b, D, None, None, 0, None, 0, C, 1)
    ref = X.double() @ torch.cat(W).double().t() + torch.cat(b).double()
    assert _err(C, ref) < 1e-5
    C1 = torch.empty_like(C)
    _call(X, D, None, M, None, D, W, D, D, 0, 4 * D, b, D, None, None, 0, None, 0, C1, 1, mode=1)
    assert 1e-5 < _err(C1, ref) < 5e-3                      # single-pass TF32 is the coarser mode


def test_input_gradient_form_split_k_residual_and_device_count():
    g = torch.Generator().manual_seed(1)
    N, D = 900, 128
    n_true = 611
    DQ = torch.randn(N, 4 * D, generator=g).to(DEV)
    G = torch.randn(N, D, generator=g).to(DEV)
    W = [torch.randn(D, D, generator=g).to(DEV) * 0.1 for _ in range(4)]
    nd = torch.tensor([n_true], dtype=torch.int32, device=DEV)
    for split in (1, 4, 16):
        C = torch.zeros(N, D, device=DEV)
        _call(DQ, 4 * D, None, N, nd, 4 * D, W, D, D, 1, D, [], 0, G, None, 0, None, 0, C, split)
        ref = G.double() + DQ.double() @ torch.cat(W).double()
        assert _err(C[:n_true], ref[:n_true]) < 1e-5, split
        assert float(C[n_true:].abs().max()) == 0.0         # rows beyond the device-side count are not touched


def test_gathered_rows_with_node_feature_tail():
    g = torch.Generator().manual_seed(2)
    nodes, D, Fn, M = 5000, 128, 1, 644
    mem = torch.randn(nodes, D, generator=g).to(DEV)
    x = torch.rand(nodes, Fn, generator=g).to(DEV)
    n_id = torch.randperm(nodes, generator=g)[:M].sort().values.to(DEV)
    Wenc = (torch.randn(D, D + Fn, generator=g) * 0.1).to(DEV)
    benc = torch.randn(D, generator=g).to(DEV)
    C = torch.empty(M, D, device=DEV)
    _call(mem, D, n_id, M, None, D, [Wenc], D + Fn, D, 0, D, [benc], D, None, x, Fn, Wenc.data_ptr() + 4 * D, D + Fn, C, 1)
    z0 = torch.cat([mem[n_id], x[n_id]], 1).double()
    ref = z0 @ Wenc.double().t() + benc.double()
    assert _err(C, ref) < 1e-5


def test_unsupported_shapes_are_refused():
    from ctan_b200._lib import CtanLibraryError
    X = torch.randn(8, 53, device=DEV)
    W = torch.randn(53, 53, device=DEV)
    C = torch.empty(8, 53, device=DEV)
    with pytest.raises(CtanLibraryError):
        _call(X, 53, None, 8, None, 53, [W], 53, 53, 0, 53, [], 0, None, None, 0, None, 0, C, 1)
