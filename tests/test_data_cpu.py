"""Host-side input path: collate_fn against the reference's pad-and-stack semantics (trainer.py:74-95)."""
import torch
import torch.nn.functional as F

from osufusion_b200.data import DevicePrefetcher, collate_fn


def _reference_collate(batch):
    """Straight restatement of trainer.py:74-95 (F.pad each item, then stack)."""
    n_max = max(x.shape[1] for x, _, _ in batch)
    xs = [F.pad(x, (0, n_max - x.shape[1]), mode="constant", value=-1.0) for x, _, _ in batch]
    as_ = [F.pad(a, (0, n_max - a.shape[1]), mode="constant", value=-23.0) for _, a, _ in batch]
    return torch.stack(xs), torch.stack(as_), torch.stack([c for _, _, c in batch]), torch.tensor([x.shape[1] for x, _, _ in batch])


def test_collate_matches_reference_semantics():
    g = torch.Generator().manual_seed(0)
    batch = [(torch.randn(6, n, generator=g), torch.randn(96, n, generator=g), torch.randn(5, generator=g)) for n in (37, 128, 1, 100)]
    out = collate_fn(batch)
    ref = _reference_collate(batch)
    assert all(torch.equal(o, r) for o, r in zip(out, ref))
    assert out[0].shape == (4, 6, 128) and out[3].tolist() == [37, 128, 1, 100] and out[3].dtype == torch.int64


def test_collate_equal_lengths_and_prefetcher_on_cpu():
    g = torch.Generator().manual_seed(1)
    batches = [collate_fn([(torch.randn(6, 16, generator=g), torch.randn(96, 16, generator=g), torch.randn(5, generator=g))
                           for _ in range(2)]) for _ in range(3)]
    assert all((b[3] == 16).all() for b in batches)
    got = list(DevicePrefetcher(batches, torch.device("cpu")))
    assert len(got) == 3 and all(torch.equal(a, b) for x, y in zip(got, batches) for a, b in zip(x, y))
    assert list(DevicePrefetcher([], torch.device("cpu"))) == []
