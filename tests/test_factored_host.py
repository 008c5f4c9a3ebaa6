"""CPU checks of the factored input format: ``FactoredFeatures`` describes exactly the
matrix the reference's preparers build with ``scipy.sparse.hstack`` (``coat/_preparer.py:154-170``,
``kuairec/_feature.py:201-207``)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden, golden_csr
from rfm_b200.factored import FactoredFeatures
from rfm_b200.synth import factored_from_tables, make_coat_shaped, make_kuairec_shaped


def _same_csr(A, B):
    A, B = A.tocsr(), B.tocsr()
    assert A.shape == B.shape
    np.testing.assert_array_equal(A.indptr, B.indptr)
    np.testing.assert_array_equal(A.indices, B.indices)
    np.testing.assert_array_equal(A.data, B.data)


def kuairec_small_log():
    return make_kuairec_shaped(seed=2025, n_users=300, n_items=400, n_train=6000, n_val=600, eval_users=60,
                               eval_items=200, eval_rows_per_user=20)


@pytest.mark.parametrize("which", ["coat", "kuairec"])
def test_factored_form_is_the_stacked_matrix_of_the_goldens(which):
    """The synthetic generators are deterministic, so the factored form can be rebuilt next to the golden inputs:
    it must materialise to the very CSR the unmodified reference was run on."""
    log = make_coat_shaped(seed=2024) if which == "coat" else kuairec_small_log()
    g = load_golden("coat_fm_ips_alpha01" if which == "coat" else "kuairec_small_fm_ips")
    for split, d in (("train", log.fm_train), ("val", log.fm_val)):
        ff = factored_from_tables(log.tables, d["users"], d["items"], d["ctx"])
        _same_csr(ff.tocsr(), golden_csr(g, split))
        assert ff.shape == golden_csr(g, split).shape
        if split == "train":                                  # the point of the format (the tables amortise over the rows)
            X = golden_csr(g, split)
            assert ff.nbytes * 4 < X.data.nbytes + X.indices.nbytes + X.indptr.nbytes
    t = log.test_frame
    ff = factored_from_tables(log.tables, t["user"], t["item"], None if which == "coat" else np.zeros(t["user"].size))
    ref = golden_csr(g, "test")
    ref.eliminate_zeros()        # the synthetic test rows store the context value 0.0 explicitly; scipy's hstack would not
    _same_csr(ff.tocsr(), ref)


def test_hstack_of_the_reference_preparer_expression():
    """coat/_preparer.py:154-170 verbatim on random tables vs the factored description of the same blocks."""
    rng = np.random.default_rng(0)
    n_users, n_items, n = 13, 17, 200
    onehot_user_ids, onehot_item_ids = sp.identity(n_users, format="csr"), sp.identity(n_items, format="csr")
    user_features = sp.random(n_users, 5, density=0.5, format="csr", random_state=1)
    item_features = sp.random(n_items, 7, density=0.3, format="csr", random_state=2)
    user_ids, item_ids = rng.integers(0, n_users, n), rng.integers(0, n_items, n)
    ref = sp.hstack([onehot_user_ids[user_ids], user_features[user_ids], onehot_item_ids[item_ids],
                     item_features[item_ids]]).tocsr()
    ff = FactoredFeatures([("id", "user", n_users), ("table", "user", user_features), ("id", "item", n_items),
                           ("table", "item", item_features)], user_ids, item_ids)
    _same_csr(ff.tocsr(), ref)
    _same_csr(ff[5:50].tocsr(), ref[5:50])
    mask = rng.random(n) < 0.3
    _same_csr(ff[mask].tocsr(), ref[mask])
    assert len(ff) == n and ff.shape == ref.shape


def test_bad_descriptions_raise():
    with pytest.raises(ValueError):
        FactoredFeatures([("id", "user", 3)], [0, 1], [0])
    with pytest.raises(ValueError):
        FactoredFeatures([("id", "nobody", 3)], [0], [0])
    with pytest.raises(ValueError):
        FactoredFeatures([("ctx", np.zeros(3))], [0], [0])
    with pytest.raises(ValueError):
        FactoredFeatures([("what", 1)], [0], [0])
    with pytest.raises(ValueError):
        FactoredFeatures([("id", "user", 2)] * 7, [0], [0])
