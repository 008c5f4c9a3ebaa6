"""ORACLE (test infrastructure, not product code) -- NumPy specification of the two-level FM step.

Only ``tests/`` may import this module. The shipped package never does.

Parity status: PINNED through ``fm_oracle`` -- this file restates ``fm_oracle.fm_grad`` / ``fm_predict`` (which are
held to the goldens of the unmodified reference) for rows given in FACTORED form, computing every sum per entity
first, exactly as ``csrc/two_level.cuh`` does; ``tests/test_two_level_host.py`` checks it against ``fm_oracle`` on the
stacked matrix at 1e-12. It is the CPU statement of the algebra the CUDA path relies on:

A factored row is ``x_t = [blocks keyed by user(t) | blocks keyed by item(t) | context values of t]`` (the reference
stacks ``onehot[user], user_table[user], onehot[item], item_table[item]``: ``utils/dataloader/coat/_preparer.py:154-170``,
``kuairec/_feature.py:169-209``). With ``E`` the (n_entities x n_features) matrix whose row v is the user-keyed (or
item-keyed) part of a row of that entity ("entity list"), and virtual columns ``[users | items | context columns]``:

* entity forward  ``A = E V``, ``alpha = E w``, ``q = (E∘E) (V∘V).sum(1)``                     (per ENTITY, not per row)
* row pass        ``s_t = A[u] + A[U+i] + sum_c x_tc v_c``; the logit of ``src/fm.py:125-132`` with ``alpha``/``q`` standing in
* level 1         ``R_v = sum_{t: v in row t} e_t x_tv s_t``,  ``a_v = sum e_t x_tv``,  ``c_v = sum e_t x_tv^2``
* level 2         ``grad v_j = sum_v E[v, j] R_v - v_j sum_v E[v, j]^2 c_v``,  ``grad w_j = sum_v E[v, j] a_v``
  (``src/fm.py:135-187`` with the sum over rows regrouped per entity)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import fm_oracle


def entity_lists(ff):
    """(E, n_users, n_items, n_ctx, ctx_cols): E is the sparse (U + I + n_ctx, n_features) matrix of entity lists over
    the stacked matrix's columns; a context column's "entity" is the column itself (one entry, 1.0)."""
    n_users = n_items = 0
    for b in ff.blocks:
        if b[0] == "id":
            n = b[2]
        elif b[0] == "table":
            n = b[2].shape[0]
        else:
            continue
        if b[1] == "user":
            n_users = n if n_users == 0 else min(n_users, n)
        else:
            n_items = n if n_items == 0 else min(n_items, n)
    parts_u, parts_i, ctx_cols, col0 = [], [], [], 0
    n_features = ff.shape[1]
    for b in ff.blocks:
        if b[0] == "id":
            n_ent = n_users if b[1] == "user" else n_items
            m = sp.csr_matrix((np.ones(n_ent), (np.arange(n_ent), col0 + np.arange(n_ent))), shape=(n_ent, n_features))
            (parts_u if b[1] == "user" else parts_i).append(m)
            col0 += b[2]
        elif b[0] == "table":
            n_ent = n_users if b[1] == "user" else n_items
            t = b[2].tocsr()[:n_ent].tocoo()
            m = sp.csr_matrix((t.data, (t.row, col0 + t.col)), shape=(n_ent, n_features))
            (parts_u if b[1] == "user" else parts_i).append(m)
            col0 += b[2].shape[1]
        else:
            ctx_cols += list(range(col0, col0 + b[1].shape[1]))
            col0 += b[1].shape[1]
    zero = lambda n: sp.csr_matrix((n, n_features))
    Eu = sum(parts_u[1:], parts_u[0]) if parts_u else zero(n_users)
    Ei = sum(parts_i[1:], parts_i[0]) if parts_i else zero(n_items)
    n_ctx = len(ctx_cols)
    Ec = sp.csr_matrix((np.ones(n_ctx), (np.arange(n_ctx), np.array(ctx_cols, dtype=np.int64))), shape=(n_ctx, n_features))
    return sp.vstack([Eu, Ei, Ec]).tocsr(), n_users, n_items, n_ctx, ctx_cols


def virtual_rows(ff, n_users, n_items, n_ctx, rows):
    """The batch as a sparse (B, U + I + n_ctx) matrix over virtual columns: 1 at the user, 1 at the item, the context
    values (exact zeros dropped, as scipy's csr_matrix(dense) drops them)."""
    B = len(rows)
    u, i = np.asarray(ff.users)[rows], np.asarray(ff.items)[rows]
    r = [np.arange(B), np.arange(B)]
    c = [u, n_users + i]
    x = [np.ones(B), np.ones(B)]
    ctx = [b[1] for b in ff.blocks if b[0] == "ctx"]
    if ctx:
        dense = np.concatenate(ctx, axis=1)[rows]
        rr, cc = np.nonzero(dense)
        r.append(rr)
        c.append(n_users + n_items + cc)
        x.append(dense[rr, cc])
    return sp.csr_matrix((np.concatenate(x), (np.concatenate(r), np.concatenate(c))),
                         shape=(B, n_users + n_items + n_ctx))


def entity_forward(E, w, V):
    """(A, alpha, q): the aggregated parameter table the row passes read."""
    return E.dot(V), E.dot(w), np.asarray(E.power(2).dot((V ** 2).sum(axis=1))).ravel()


def two_level_predict(Xv, w0, A, alpha, q):
    """src/fm.py:125-132 on virtual rows: identical to fm_oracle.fm_predict on the stacked rows."""
    S = Xv.dot(A)
    lin = Xv.dot(alpha)
    qq = np.asarray(Xv.power(2).dot(q)).ravel()
    return fm_oracle.sigmoid(float(np.asarray(w0).ravel()[0]) + lin + 0.5 * ((S ** 2).sum(axis=1) - qq)), S


def two_level_grad(ff, rows, y, ps, w0, w, V):
    """(sum e, grad w, grad V) of the batch ``rows`` -- the quantities of fm_oracle.fm_grad -- through the two levels."""
    E, n_users, n_items, n_ctx, _ = entity_lists(ff)
    Xv = virtual_rows(ff, n_users, n_items, n_ctx, rows)
    A, alpha, q = entity_forward(E, w, V)
    p, S = two_level_predict(Xv, w0, A, alpha, q)
    e = y / ps - p
    Xe = Xv.multiply(e[:, None]).tocsr()
    R = Xe.T.dot(S)                                                   # level 1: per virtual column
    a = np.asarray(Xe.sum(axis=0)).ravel()
    c = np.asarray(Xv.power(2).multiply(e[:, None]).sum(axis=0)).ravel()
    gw = E.T.dot(a)                                                   # level 2: per real column
    gV = E.T.dot(R) - np.asarray(E.power(2).T.dot(c)).ravel()[:, None] * V
    return e.sum(), np.asarray(gw).ravel(), np.asarray(gV)
