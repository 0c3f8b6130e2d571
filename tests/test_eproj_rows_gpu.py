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
