"""Drop-in surface of the loop functions (CPU): `ctan_b200.train_link` / `ctan_b200.train_sequence` /
`ctan_b200.negative_sampler` expose the reference's parameter names, order and defaults (train_link.py:24, 206, 287, 349;
train_sequence.py:16, 68; negative_sampler.py:9, 27), checked against the reference's own functions imported unmodified."""
import inspect

import pytest


def _ref():
    from oracle import ref_loops
    if not ref_loops.available():
        pytest.skip("reference files not present (neither /root/reference nor baseline/_ref/)")
    return ref_loops.load()


def _params(fn):
    fn = inspect.unwrap(fn)
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]


@pytest.mark.parametrize("name", ["tgb_train", "tgb_test", "train", "eval"])
def test_train_link_signatures(name):
    R = _ref()
    import ctan_b200  # noqa: F401
    from ctan_b200 import train_link as TL
    assert _params(getattr(TL, name)) == _params(getattr(R.train_link, name))


@pytest.mark.parametrize("name", ["train", "eval"])
def test_train_sequence_signatures(name):
    R = _ref()
    import ctan_b200  # noqa: F401
    from ctan_b200 import train_sequence as TS
    assert _params(getattr(TS, name)) == _params(getattr(R.train_sequence, name))


def test_negative_sampler_surface():
    R = _ref()
    import ctan_b200  # noqa: F401
    from ctan_b200 import negative_sampler as NS
    ref = R.negative_sampler.NegativeSampler
    mine = NS.NegativeSampler
    assert _params(mine.sample) == _params(ref.sample)
    assert _params(mine.__init__)[:6] == _params(ref.__init__)       # ours adds a trailing `device=None`
    assert NS.neg_sampler_names == R.negative_sampler.neg_sampler_names


def test_temporal_data_slicing_rule():
    """The minimal TemporalData / TemporalDataLoader shipped for PyG-less use follow PyG's rule."""
    import torch
    import ctan_b200  # noqa: F401
    from ctan_b200 import train_link as TL
    d = TL.TemporalData(src=torch.arange(10), dst=torch.arange(10) + 10, t=torch.arange(10), msg=torch.zeros(10, 3),
                        y=torch.arange(10))
    d.x = torch.ones(1, 20, 2)
    b = d[2:5]
    assert b.num_events == 3 and b.x is d.x and torch.equal(b.y, torch.tensor([2, 3, 4])) and d.num_nodes == 20
    sizes = [x.num_events for x in TL.TemporalDataLoader(d, batch_size=4)]
    assert sizes == [4, 4, 2] and len(TL.TemporalDataLoader(d, batch_size=4)) == 3
