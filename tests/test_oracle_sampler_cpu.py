"""CPU: the oracle's Philox draw (oracle/vampnet_oracle.py sample_from_logits, rng="philox") is a two-level inverse CDF —
uniform 1 picks a 128-entry vocabulary tile by mass, uniform 2 the entry inside it.  The kernels are tested for equality
with it; this pins that the definition itself samples softmax(logits / T), whatever the tile structure of the mass."""
import numpy as np
import pytest
import torch

from oracle import vampnet_oracle as vo


def _oracle():
    cfg = vo.OracleConfig(n_heads=4, n_layers=1, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=256)
    return vo.OracleVampNet(cfg, vo.make_state_dict(cfg, seed=0), "fp32")


@pytest.mark.parametrize("layout", ["spread", "one_tile", "tile_edges"])
def test_two_level_draw_samples_the_softmax(layout):
    g = torch.Generator().manual_seed(3)
    V, S, T = 1024, 30000, 0.8
    row = torch.full((V,), -40.0)
    if layout == "spread":          # mass in several tiles
        support = torch.randperm(V, generator=g)[:48]
    elif layout == "one_tile":      # all mass inside tile 5
        support = 5 * 128 + torch.randperm(128, generator=g)[:20]
    else:                           # first / last entries of tiles: the crossing logic at tile boundaries
        support = torch.tensor([0, 127, 128, 255, 256, 511, 512, 1023])
    row[support] = torch.randn(len(support), generator=g) * 1.5
    logits = row[None, None, :].expand(1, S, V).contiguous()
    tok, ptok = _oracle().sample_from_logits(logits, True, T, rng="philox", philox_key=(11, 7), step=2)
    p = torch.softmax(row.double() / T, 0)
    counts = torch.bincount(tok.flatten(), minlength=V).double()
    assert counts[p < 1e-9].sum() == 0
    keep = p * S >= 5
    chi2 = (((counts - p * S) ** 2) / (p * S))[keep].sum().item()
    dof = int(keep.sum()) - 1
    assert chi2 < dof + 5 * (2 * dof) ** 0.5, (chi2, dof)
    assert torch.allclose(ptok.flatten().double(), p[tok.flatten()], rtol=1e-4, atol=1e-9)


def test_greedy_and_degenerate_rows():
    o = _oracle()
    logits = torch.full((1, 3, 1024), -50.0)
    logits[0, 0, 700] = 3.0                       # one-hot: always that token
    logits[0, 1, :] = 0.0                         # flat: any token, probability 1/1024
    logits[0, 2, 130] = 1.0
    logits[0, 2, 131] = 1.0                       # tie inside a tile: arg-max takes the lowest index when not sampling
    tok, p = o.sample_from_logits(logits, True, 1.0, rng="philox", philox_key=(1, 2), step=0)
    assert tok[0, 0] == 700 and abs(float(p[0, 1]) - 1 / 1024) < 1e-6 and tok[0, 2] in (130, 131)
    tok, _ = o.sample_from_logits(logits, False, 1.0, rng="philox", philox_key=(1, 2), step=0)
    assert tok[0, 0] == 700 and tok[0, 2] == 130
