"""Seeded synthetic Coat-shaped and KuaiRec-shaped inputs.

The real datasets are not shipped with the reference (``.gitignore:1-4``), so every
parity test and benchmark runs on synthetic data that has the *shapes and dtypes* the
reference's data layer hands to ``fit`` / ``predict`` / the evaluators:

* FM rows: ``scipy.sparse.csr_matrix`` float64 data, int32 indices, sorted columns.
  Coat layout  ``[I_user | user_feat | I_item | item_feat]``  (reference
  ``utils/dataloader/coat/_preparer.py:154-170``); KuaiRec layout
  ``[I_user | I_item | timestamp | user_feat | video_feat]``
  (``utils/dataloader/kuairec/_feature.py:54-84,201-207``).
* MF rows: int64 ``(N, 2)`` ``[user, item]``; labels int64; pscores float64.
* Evaluator frames: columns ``user, item, label, pscore, ones_pscore``.

Everything is drawn from ``np.random.default_rng(seed)`` so it never touches the legacy
global NumPy RNG the models seed (``src/fm.py:34``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
from scipy.sparse import csr_matrix

__all__ = ["SyntheticLog", "make_coat_shaped", "make_kuairec_shaped", "csr_from_tables", "factored_from_tables"]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


@dataclass
class SyntheticLog:
    """One synthetic experiment: train/val dicts for FM and MF plus an evaluator frame."""

    n_users: int
    n_items: int
    n_features: int
    fm_train: Dict[str, object]
    fm_val: Dict[str, object]
    mf_train: Dict[str, np.ndarray]
    mf_val: Dict[str, np.ndarray]
    # held-out rows for TestEvaluator / ValEvaluator: dict of equally long 1-D arrays
    test_frame: Dict[str, np.ndarray]
    fm_test_features: Optional[csr_matrix]
    mf_test_features: np.ndarray
    # the per-entity tables the rows were assembled from (the "factored" format)
    tables: Dict[str, object]


def csr_from_tables(users, items, tables, ctx=None) -> csr_matrix:
    """Assemble FM rows for (user, item[, ctx]) triples from per-user / per-item tables.

    ``tables`` holds, per side, a ragged list of (column, value) pairs in CSR form:
    ``u_ptr/u_col/u_val`` (columns already global) and ``i_ptr/i_col/i_val``; ``ctx_col``
    is the global column of the single real-valued context feature (or None).
    Column order inside each row follows ``tables["order"]`` so the result has sorted
    indices exactly like ``scipy.sparse.hstack`` of the reference's blocks.
    """
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    n_rows = users.shape[0]
    u_ptr, i_ptr = tables["u_ptr"], tables["i_ptr"]
    ulen = (u_ptr[users + 1] - u_ptr[users]).astype(np.int64)
    ilen = (i_ptr[items + 1] - i_ptr[items]).astype(np.int64)
    has_ctx = tables.get("ctx_col") is not None
    row_len = ulen + ilen + (1 if has_ctx else 0)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(row_len, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, dtype=np.int32)
    data = np.empty(nnz, dtype=np.float64)

    def scatter(side_ptr, side_col, side_val, ids, lens, dst_start):
        # ragged gather: for each row r copy side[ids[r]] to indices[dst_start[r]: +lens[r]]
        total = int(lens.sum())
        if total == 0:
            return
        row_of = np.repeat(np.arange(n_rows, dtype=np.int64), lens)
        first = np.zeros(n_rows, dtype=np.int64)
        np.cumsum(lens[:-1], out=first[1:])
        within = np.arange(total, dtype=np.int64) - first[row_of]
        src = side_ptr[ids][row_of] + within
        dst = dst_start[row_of] + within
        indices[dst] = side_col[src]
        data[dst] = side_val[src]

    # tables are built so that, per row, the global columns come out ascending when the
    # blocks are written in tables["order"]; each block is itself ascending.
    cursor = indptr[:-1].copy()
    for block in tables["order"]:
        if block == "user":
            scatter(u_ptr, tables["u_col"], tables["u_val"], users, ulen, cursor)
            cursor = cursor + ulen
        elif block == "item":
            scatter(i_ptr, tables["i_col"], tables["i_val"], items, ilen, cursor)
            cursor = cursor + ilen
        elif block == "ctx":
            indices[cursor] = tables["ctx_col"]
            data[cursor] = 0.0 if ctx is None else np.asarray(ctx, dtype=np.float64)
            cursor = cursor + 1
        else:  # pragma: no cover
            raise ValueError(block)
    if nnz < 2**31 - 1:
        indptr = indptr.astype(np.int32)
    X = csr_matrix((data, indices, indptr), shape=(n_rows, tables["n_features"]))
    X.has_sorted_indices = True
    return X


def factored_from_tables(tables, users, items, ctx=None):
    """The same rows as ``csr_from_tables`` (after its column sort) in the factored form (SURVEY.md section 8 f3,
    ``rfm_b200.factored.FactoredFeatures``): the blocks the reference's preparers stack -- Coat
    ``[I_user | user_feat | I_item | item_feat]`` (``coat/_preparer.py:154-170``), KuaiRec
    ``[I_user | I_item | ctx | user_feat | item_feat]`` (``kuairec/_feature.py:201-207``) -- plus the id pairs."""
    from .factored import FactoredFeatures
    n_users, n_items = tables["n_users"], tables["n_items"]

    def side_table(ptr, col, val, base, width):
        # every table row starts with the entity's one-hot id entry; the rest are its side features (global columns)
        n = ptr.shape[0] - 1
        keep = np.ones(col.shape[0], dtype=bool)
        keep[ptr[:-1]] = False
        new_ptr = ptr - np.arange(n + 1)
        return csr_matrix((val[keep], (col[keep] - base).astype(np.int32), new_ptr), shape=(n, width))

    ut = side_table(tables["u_ptr"], tables["u_col"], tables["u_val"], tables["uf_base"], tables["n_uf"])
    it = side_table(tables["i_ptr"], tables["i_col"], tables["i_val"], tables["if_base"], tables["n_if"])
    if tables.get("ctx_col") is None:
        blocks = [("id", "user", n_users), ("table", "user", ut), ("id", "item", n_items), ("table", "item", it)]
    else:
        c = np.zeros(len(users)) if ctx is None else np.asarray(ctx, dtype=np.float64)
        blocks = [("id", "user", n_users), ("id", "item", n_items), ("ctx", c), ("table", "user", ut),
                  ("table", "item", it)]
    return FactoredFeatures(blocks, np.asarray(users), np.asarray(items))


def _ragged(rows_cols, rows_vals):
    ptr = np.zeros(len(rows_cols) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in rows_cols], out=ptr[1:])
    col = np.concatenate(rows_cols).astype(np.int32) if len(rows_cols) else np.zeros(0, np.int32)
    val = np.concatenate(rows_vals).astype(np.float64) if len(rows_vals) else np.zeros(0)
    return ptr, col, val


def _popularity(n_items, rng):
    rank = rng.permutation(n_items)
    pop = (rank + 10.0) ** -0.8
    return pop / pop.sum()


def _negative_sample(labels, seed):
    """positives first, then an equal number of negatives chosen by the legacy global RNG
    (reference ``utils/dataloader/kuairec/_preparer.py:104-115``). Uses a private
    RandomState so the global stream is untouched."""
    pos = np.flatnonzero(labels == 1)
    neg = np.flatnonzero(labels == 0)
    rs = np.random.RandomState(seed)
    neg = rs.permutation(neg)[: pos.shape[0]]
    return np.concatenate([pos, neg])


def make_coat_shaped(seed: int = 2024, n_users: int = 290, n_items: int = 300,
                     n_rated: int = 24, n_test: int = 16, pow_used: float = 0.1,
                     val_ratio: float = 0.2, negative_sampling: bool = True) -> SyntheticLog:
    """C1/C2 shape: 290 users x 300 items, 14 user + 33 item binary side features, n=637."""
    rng = np.random.default_rng(seed)
    u_groups, i_groups = (2, 6, 3, 3), (2, 16, 13, 2)
    n_uf, n_if = sum(u_groups), sum(i_groups)
    n_features = n_users + n_uf + n_items + n_if
    # column blocks: [I_user | user_feat | I_item | item_feat]
    uf_base, item_base = n_users, n_users + n_uf
    if_base = item_base + n_items

    def one_hot_groups(n, groups, base):
        cols = []
        off = base
        for g in groups:
            cols.append(off + rng.integers(0, g, size=n))
            off += g
        return np.stack(cols, axis=1)

    ufe = one_hot_groups(n_users, u_groups, uf_base)
    ife = one_hot_groups(n_items, i_groups, if_base)
    u_cols = [np.concatenate([[u], ufe[u]]) for u in range(n_users)]
    i_cols = [np.concatenate([[item_base + i], ife[i]]) for i in range(n_items)]
    u_ptr, u_col, u_val = _ragged(u_cols, [np.ones(len(c)) for c in u_cols])
    i_ptr, i_col, i_val = _ragged(i_cols, [np.ones(len(c)) for c in i_cols])
    tables = dict(u_ptr=u_ptr, u_col=u_col, u_val=u_val, i_ptr=i_ptr, i_col=i_col, i_val=i_val,
                  ctx_col=None, order=("user", "item"), n_features=n_features, n_users=n_users, n_items=n_items,
                  uf_base=uf_base, n_uf=n_uf, if_base=if_base, n_if=n_if)

    pop = _popularity(n_items, rng)
    hidden_p = rng.normal(size=(n_users, 8)) * 0.6
    hidden_q = rng.normal(size=(n_items, 8)) * 0.6
    theta = 0.11 + 0.89 * (pop / pop.max())            # exposure in [0.11, 1]
    users = np.repeat(np.arange(n_users), n_rated)
    items = np.concatenate([rng.choice(n_items, size=n_rated, replace=False, p=pop)
                            for _ in range(n_users)])
    gamma = _sigmoid((hidden_p[users] * hidden_q[items]).sum(1) + rng.normal(size=users.size) * 0.3)
    labels = (rng.random(users.size) < gamma).astype(np.int64)
    pscores = theta[items] ** pow_used
    tables["item_pscore"] = theta ** pow_used          # the per-item table the per-row pscores are gathered from

    perm = rng.permutation(users.size)
    n_val = int(round(users.size * val_ratio))
    val_sel, tr_sel = perm[:n_val], perm[n_val:]

    def subset(sel, ns_seed):
        if negative_sampling:
            sel = sel[_negative_sample(labels[sel], ns_seed)]
        return sel

    tr_sel, val_sel = subset(tr_sel, 12345), subset(val_sel, 12346)

    def dicts(sel):
        X = csr_from_tables(users[sel], items[sel], tables)
        fm = {"features": X, "labels": labels[sel].copy(), "pscores": pscores[sel].copy(),
              "users": users[sel].copy(), "items": items[sel].copy(), "ctx": None}
        mf = {"features": np.stack([users[sel], items[sel]], axis=1).astype(np.int64),
              "labels": labels[sel].copy(), "pscores": pscores[sel].copy()}
        return fm, mf

    fm_train, mf_train = dicts(tr_sel)
    fm_val, mf_val = dicts(val_sel)

    t_users = np.repeat(np.arange(n_users), n_test)
    t_items = np.concatenate([rng.choice(n_items, size=n_test, replace=False) for _ in range(n_users)])
    t_gamma = _sigmoid((hidden_p[t_users] * hidden_q[t_items]).sum(1))
    t_labels = (rng.random(t_users.size) < t_gamma).astype(np.int64)
    order = rng.permutation(t_users.size)               # interleave users like a real log
    t_users, t_items, t_labels = t_users[order], t_items[order], t_labels[order]
    test_frame = dict(user=t_users.astype(np.int64), item=t_items.astype(np.int64), label=t_labels,
                      pscore=theta[t_items].copy(), ones_pscore=np.ones(t_users.size))
    return SyntheticLog(n_users, n_items, n_features, fm_train, fm_val, mf_train, mf_val, test_frame,
                        csr_from_tables(t_users, t_items, tables),
                        np.stack([t_users, t_items], axis=1).astype(np.int64), tables)


def make_kuairec_shaped(seed: int = 2024, n_users: int = 7176, n_items: int = 10728,
                        n_train: int = 12_000_000, n_val: int = 2000,
                        eval_users: int = 1411, eval_items: int = 3327, eval_rows_per_user: int = 46,
                        pow_used: float = 0.5, exposure_bias: float = 3.0,
                        build_mf: bool = True, build_eval: bool = True) -> SyntheticLog:
    """C3/C4 shape: one-hot user + item ids, 1 real-valued context column, 7 user one-hot
    groups (~90 columns), 4 real-valued + 1..4-of-31 multi-hot item columns; n ~ 18k, m ~ 16.5.

    Fully vectorised so the 12 M-row configuration builds in seconds on the host.
    """
    rng = np.random.default_rng(seed)
    u_groups = (9, 8, 30, 12, 4, 15, 12)                # ~90 user one-hot columns
    n_uf, n_real, n_cat = sum(u_groups), 4, 31
    ctx_col = n_users + n_items
    uf_base = ctx_col + 1
    if_base = uf_base + n_uf
    n_features = if_base + n_real + n_cat

    ufe = np.empty((n_users, 1 + len(u_groups)), dtype=np.int64)
    ufe[:, 0] = np.arange(n_users)
    off = uf_base
    for g_idx, g in enumerate(u_groups):
        ufe[:, 1 + g_idx] = off + rng.integers(0, g, size=n_users)
        off += g
    u_ptr = np.arange(n_users + 1, dtype=np.int64) * ufe.shape[1]
    u_col, u_val = ufe.reshape(-1).astype(np.int32), np.ones(ufe.size)

    n_cats = rng.integers(1, 5, size=n_items)
    cat_choice = np.argsort(rng.random((n_items, n_cat)), axis=1)[:, :4]
    cat_choice = np.where(np.arange(4)[None, :] < n_cats[:, None], cat_choice, n_cat + 7)
    cat_choice.sort(axis=1)                             # unused slots (sentinel) sort last
    real_vals = rng.normal(size=(n_items, n_real))
    i_len = 1 + n_real + n_cats
    i_ptr = np.zeros(n_items + 1, dtype=np.int64)
    np.cumsum(i_len, out=i_ptr[1:])
    i_col = np.empty(int(i_ptr[-1]), dtype=np.int32)
    i_val = np.empty(int(i_ptr[-1]), dtype=np.float64)
    base = i_ptr[:-1]
    i_col[base] = n_users + np.arange(n_items)
    i_val[base] = 1.0
    for r in range(n_real):
        i_col[base + 1 + r] = if_base + r
        i_val[base + 1 + r] = real_vals[:, r]
    for c in range(4):
        sel = n_cats > c
        i_col[base[sel] + 1 + n_real + c] = if_base + n_real + cat_choice[sel, c]
        i_val[base[sel] + 1 + n_real + c] = 1.0
    # global column order inside a row: user id < item id < ctx < user feats < item feats.
    # The item table's first entry (the one-hot id) sorts before ctx, the rest after, so the
    # item side is split into two blocks to keep indices ascending.
    tables = dict(u_ptr=u_ptr, u_col=u_col, u_val=u_val, i_ptr=i_ptr, i_col=i_col, i_val=i_val,
                  ctx_col=ctx_col, n_features=n_features, order=("user", "item", "ctx"), n_users=n_users,
                  n_items=n_items, uf_base=uf_base, n_uf=n_uf, if_base=if_base, n_if=n_real + n_cat)

    pop = _popularity(n_items, rng)
    z = rng.normal(size=n_items)
    theta = np.maximum(_sigmoid(3.0 * z - 1.0) ** exposure_bias, 0.1)   # kuairec/_click.py:193-202
    tables["item_pscore"] = theta ** pow_used          # the per-item table the per-row pscores are gathered from
    hidden_p = rng.normal(size=(n_users, 8)) * 0.5
    hidden_q = rng.normal(size=(n_items, 8)) * 0.5
    activity = rng.lognormal(sigma=0.6, size=n_users)
    activity /= activity.sum()

    def draw(n_rows, item_pool=None, user_pool=None):
        u = rng.choice(n_users if user_pool is None else user_pool, size=n_rows,
                       p=activity if user_pool is None else None)
        if item_pool is None:
            i = rng.choice(n_items, size=n_rows, p=pop)
        else:
            i = rng.choice(item_pool, size=n_rows)
        g = _sigmoid(np.einsum("ij,ij->i", hidden_p[u], hidden_q[i]) + rng.normal(size=n_rows) * 0.3)
        r = (rng.random(n_rows) < g)
        o = (rng.random(n_rows) < theta[i])
        return u.astype(np.int64), i.astype(np.int64), r, o

    def sorted_rows(users, items, ctx):
        # rows are [user block | item id | ctx | user feats | item feats]; build with the
        # generic assembler, then restore ascending column order per row with one argsort of
        # a (row, col) key -- cheap and keeps this generator obviously correct.
        X = csr_from_tables(users, items, tables, ctx)
        X.has_sorted_indices = False
        X.sort_indices()
        return X

    def dicts(n_rows):
        u, i, r, o = draw(n_rows)
        y = (r & o).astype(np.int64)
        ps = theta[i] ** pow_used
        ctx = rng.normal(size=n_rows)
        fm = {"features": sorted_rows(u, i, ctx), "labels": y, "pscores": ps, "users": u, "items": i, "ctx": ctx}
        mf = ({"features": np.stack([u, i], axis=1), "labels": y.copy(), "pscores": ps.copy()}
              if build_mf else None)
        return fm, mf

    fm_train, mf_train = dicts(n_train)
    fm_val, mf_val = dicts(n_val)

    if build_eval:
        e_users = np.repeat(rng.choice(n_users, size=eval_users, replace=False), eval_rows_per_user)
        pool = rng.choice(n_items, size=eval_items, replace=False)
        e_items = rng.choice(pool, size=e_users.size)
        g = _sigmoid(np.einsum("ij,ij->i", hidden_p[e_users], hidden_q[e_items]))
        e_labels = (rng.random(e_users.size) < g).astype(np.int64)
        order = rng.permutation(e_users.size)
        e_users, e_items, e_labels = e_users[order], e_items[order], e_labels[order]
        test_frame = dict(user=e_users.astype(np.int64), item=e_items.astype(np.int64), label=e_labels,
                          pscore=theta[e_items].copy(), ones_pscore=np.ones(e_users.size))
        fm_test = sorted_rows(e_users, e_items, np.zeros(e_users.size))
        mf_test = np.stack([e_users, e_items], axis=1).astype(np.int64)
    else:
        test_frame, fm_test, mf_test = {}, None, np.zeros((0, 2), np.int64)
    return SyntheticLog(n_users, n_items, n_features, fm_train, fm_val, mf_train, mf_val, test_frame,
                        fm_test, mf_test, tables)
