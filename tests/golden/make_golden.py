"""
Writes tests/golden/reference_fixtures.json and tests/golden/kat.json.

The reference cannot be imported here (every hot-path module imports TensorFlow at top level and
TensorFlow is not installed), so the vectors are of two kinds:

1. reference_fixtures.json -- literal inputs / expected outputs transcribed from the reference's own
   tests (file:line given per entry).  Nothing is computed; this is what pins the oracle.
2. kat.json -- known-answer vectors for the pieces no reference test covers.  They are computed here
   with straight-line float64 numpy that does NOT import oracle/ (an independent restatement), and
   cross-checked against the constants printed in SURVEY.md section 9.

Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_fixtures():
    return {
        "logq": {
            "source": "tests/test_layers.py:8-13 (logits), :16-18 (ids), :21-23 (probs), :28-36 (expected)",
            "logits": [[1.0, -1.5, 2.5], [-1.0, -2.5, 1.5], [2.5, -1.5, -1.0]],
            "candidate_ids": ["id1", "id2", "id3"],
            "candidate_prob_lookup": {"id1": 0.3, "id2": 0.2, "id3": 0.5},
            "expected": [
                [2.2039728043, 0.1094379124, 3.1931471806],
                [0.2039728043, -0.8905620876, 2.1931471806],
                [3.7039728043, 0.1094379124, -0.3068528194],
            ],
        },
        "brute_force": {
            "source": "tests/test_indices.py:63-80 (query table), :83-102 (candidates), :110-129 (queries, expected)",
            "query_vocab": ["query_1", "query_2", "query_3"],
            # row 0 is the OOV row of StringLookup (MockEmbeddingModel, tests/test_indices.py:24-47)
            "query_table": [[1.0, 1.0], [0.5, -1.0], [1.0, -0.5], [-1.0, -0.5]],
            "candidate_ids": ["candidate_1", "candidate_2", "candidate_3", "candidate_4", "candidate_5"],
            "candidate_embeddings": [[2.0, -1.5], [-1.5, 3.0], [-0.5, -1.0], [1.0, -1.5], [-2.0, -1.5]],
            "candidate_batch": 1,
            "queries": ["query_1", "query_2", "query_3", "query_4", "query_1"],
            "k": 2,
            "expected": [
                ["candidate_1", "candidate_4"],
                ["candidate_1", "candidate_4"],
                ["candidate_5", "candidate_3"],
                ["candidate_2", "candidate_1"],
                ["candidate_1", "candidate_4"],
            ],
        },
        "recall": {
            "source": "tests/test_recall.py:8-40 (test ds, batch 2), :43-76 (static index k=5), :84-95 (expected)",
            "query_ids": ["query1", "query2", "query3", "query4", "query5"],
            "true_candidate_ids": ["id1", "id7", "id2", "id2", "id10"],
            "batch_size": 2,
            "static_candidates": ["id1", "id2", "id3", "id4", "id5", "id6", "id7", "id8", "id9", "id10"],
            "static_k": 5,
            "ks": [1, 2, 5],
            "expected": {"1": 0.2, "2": 0.6, "5": 0.6},
        },
    }


def kat_a():
    """One full train step, B=3, e=E=2, one id feature per tower, no hidden layer, duplicate
    candidate in the batch (SURVEY.md section 9 KAT-A)."""
    f8 = np.float64
    Tq = np.array([[0, 0], [.5, -.25], [.25, .5], [-.5, .75]], f8)
    Tc = np.array([[0, 0], [.5, .5], [-.25, 1]], f8)
    Wq = np.array([[1, .5], [-.5, 1]], f8); bq = np.array([0, .25], f8)
    Wc = np.array([[.5, 1], [1, -.5]], f8); bc = np.array([.25, 0], f8)
    qid = np.array([1, 2, 3]); cid = np.array([1, 2, 1])
    p_row = np.array([1.0, .75, .25], f8)          # sampling prob per candidate TABLE ROW
    xq, xc = Tq[qid], Tc[cid]
    Q = np.maximum(xq @ Wq + bq, 0); C = np.maximum(xc @ Wc + bc, 0)
    S = Q @ C.T
    Z = S - np.log(p_row[cid])[None, :]
    m = Z.max(1, keepdims=True); e = np.exp(Z - m); s = e.sum(1, keepdims=True)
    lse = (m + np.log(s)).ravel()
    loss = float((lse - np.diag(Z)).sum())
    dZ = e / s - np.eye(3)
    dQ = dZ @ C; dC = dZ.T @ Q
    dpq = dQ * (Q > 0); dpc = dC * (C > 0)
    dWq = xq.T @ dpq; dbq = dpq.sum(0); dxq = dpq @ Wq.T
    dWc = xc.T @ dpc; dbc = dpc.sum(0); dxc = dpc @ Wc.T
    lr, eps, acc0 = 0.05, 1e-7, 0.1

    def sparse_adagrad(T, ids, dx):
        T = T.copy(); A = np.full_like(T, acc0)
        for r in np.unique(ids):
            g = dx[ids == r].sum(0)
            A[r] += g * g
            T[r] -= lr * g / (np.sqrt(A[r]) + eps)
        return T, A

    Tc2, Ac2 = sparse_adagrad(Tc, cid, dxc)
    Tq2, Aq2 = sparse_adagrad(Tq, qid, dxq)

    def dense_adagrad(W, g):
        A = np.full_like(W, acc0) + g * g
        return W - lr * g / (np.sqrt(A) + eps), A

    Wq2, _ = dense_adagrad(Wq, dWq); bq2, _ = dense_adagrad(bq, dbq)
    Wc2, _ = dense_adagrad(Wc, dWc); bc2, _ = dense_adagrad(bc, dbc)
    # cross-check against the constants printed in SURVEY.md section 9
    assert abs(loss - 3.724187101137391) < 1e-12
    assert np.allclose(dZ[0], [-0.80187203686, 0.60374407371, 0.19812796314], atol=1e-10)
    assert np.allclose(Tc2[1], [0.53749864515, 0.53394550928], atol=1e-10)
    assert np.allclose(Ac2[2], [0.13559637759, 0.24238551037], atol=1e-10)
    L = lambda a: np.asarray(a).tolist()
    return {
        "source": "SURVEY.md section 9 KAT-A (recomputed in float64 by tests/golden/make_golden.py)",
        "Tq": L(Tq), "Tc": L(Tc), "Wq": L(Wq), "bq": L(bq), "Wc": L(Wc), "bc": L(bc),
        "query_ids": L(qid), "candidate_ids": L(cid), "p_row": L(p_row),
        "lr": lr, "eps": eps, "acc0": acc0,
        "Q": L(Q), "C": L(C), "S": L(S), "Z": L(Z), "loss": loss, "dZ": L(dZ), "dQ": L(dQ), "dC": L(dC),
        "dWq": L(dWq), "dbq": L(dbq), "dWc": L(dWc), "dbc": L(dbc), "dxq": L(dxq), "dxc": L(dxc),
        "Tq_after": L(Tq2), "Tc_after": L(Tc2), "acc_q_after": L(Aq2), "acc_c_after": L(Ac2),
        "Wq_after": L(Wq2), "bq_after": L(bq2), "Wc_after": L(Wc2), "bc_after": L(bc2),
    }


def kat_b(fix):
    z = np.array(fix["logq"]["logits"], np.float64) - np.log(
        np.array([fix["logq"]["candidate_prob_lookup"][i] for i in fix["logq"]["candidate_ids"]], np.float64))[None, :]
    m = z.max(1, keepdims=True)
    lse = (m + np.log(np.exp(z - m).sum(1, keepdims=True))).ravel()
    loss = float((lse - np.diag(z)).sum())
    # SURVEY.md prints 8.645022182948873 (agrees to 3e-9 relative; it was evaluated in lower precision)
    assert abs(loss - 8.645022182948873) < 1e-7
    return {"source": "SURVEY.md section 9 KAT-B: CE-SUM on the reference logQ fixture (float64)", "loss": loss}


def kat_c():
    return {
        "source": "SURVEY.md section 9 KAT-C: tf.math.top_k tie rule (lower index first)",
        "scores": [[1, 3, 3, 0, 3, 2], [5, 5, 5, 5, 5, 5]], "k": 3, "indices": [[1, 2, 4], [0, 1, 2]],
    }


def main():
    fix = reference_fixtures()
    with open(os.path.join(HERE, "reference_fixtures.json"), "w") as f:
        json.dump(fix, f, indent=1)
    kat = {"A": kat_a(), "B": kat_b(fix), "C": kat_c()}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote reference_fixtures.json, kat.json")


if __name__ == "__main__":
    main()
