"""CPU checks of the click generator's specification (oracle/clicks_oracle.py)."""
import numpy as np

from oracle import clicks_oracle as co


def test_philox4x32_10_known_answers():
    """Random123's published known-answer vectors for Philox4x32-10 (counter words, key words -> output words)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        index = np.array([ctr[0] | (ctr[1] << 32)], dtype=np.uint64)
        got = co.philox4x32(key[0] | (key[1] << 32), index, ctr[2], ctr[3])[0]
        assert tuple(int(x) for x in got) == want


def test_exposure_and_relevance_follow_the_reference_formulas():
    """kuairec/_click.py:168-171 and :193-202 written out with pandas, as the reference does."""
    import pandas as pd
    rng = np.random.default_rng(3)
    counts = pd.Series(rng.integers(1, 5000, 400).astype(np.float64))
    z = (counts - counts.mean()) / counts.std()
    ref = np.maximum(z.apply(lambda x: 1 / (1 + np.exp(-(3.0 * x + -1.0)))).values ** 3.0, 0.1)
    np.testing.assert_allclose(co.exposure(counts.values, 3.0, 0.1), ref, rtol=1e-15)
    w = rng.lognormal(size=100)
    np.testing.assert_array_equal(co.relevance(w), np.clip(w / 2.0, 0, 1))


def test_generated_log_statistics_and_shard_consistency():
    from rfm_b200.clicks import ClickModel
    m = ClickModel(500, 800, seed=11)
    full = co.generate(m.seed, 0, 30000, m.user_cdf, m.item_cdf, m.item_exposure, m.pow_used, n_ctx=2)
    part = co.generate(m.seed, 12000, 5000, m.user_cdf, m.item_cdf, m.item_exposure, m.pow_used, n_ctx=2)
    for key in ("users", "items", "ctx", "labels", "relevance", "targets"):
        np.testing.assert_array_equal(part[key], full[key][12000:17000])
    # draws follow the distributions: users ~ activity, items ~ popularity, O ~ Be(theta), R ~ Be(gamma), Y = O R
    np.testing.assert_allclose(np.bincount(full["users"], minlength=500) / 30000, m.user_prob, atol=0.004)
    np.testing.assert_allclose(np.bincount(full["items"], minlength=800) / 30000, m.item_prob, atol=0.004)
    assert abs((full["u_exposure"] < full["theta"]).mean() - full["theta"].mean()) < 0.01
    assert abs(full["relevance"].mean() - full["gamma"].mean()) < 0.01
    np.testing.assert_array_equal(full["labels"], (full["u_exposure"] < full["theta"]) * full["relevance"])
    assert abs(full["ctx"].mean()) < 0.02 and abs(full["ctx"].std() - 1) < 0.02
    assert full["theta"].min() >= 0.1 and full["pscores"].max() <= 1.0
    np.testing.assert_array_equal(full["targets"], full["labels"] / full["pscores"])
