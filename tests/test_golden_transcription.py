"""The reference fixtures in tests/golden/reference_fixtures.json are transcriptions of literals in the reference's own tests.
When the reference checkout is present (this container; not the GPU box) the transcription is verified against the source text:
every list / dict we copied must appear as a literal in the cited reference test file.  Parsed with ``ast`` -- the reference's
tests import TensorFlow and cannot be imported here."""
import ast
import json
import os
import re

import pytest

REF = "/root/reference/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _literals(path):
    """Every pure-literal list / dict / tuple node of a python file, evaluated."""
    with open(path) as f:
        tree = ast.parse(f.read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, (ast.List, ast.Dict, ast.Tuple)):
            try:
                out.append(ast.literal_eval(node))
            except (ValueError, SyntaxError):
                pass
    return out


def _flat(x):
    return [y for v in x for y in _flat(v)] if isinstance(x, (list, tuple)) else [x]


@pytest.fixture(scope="module")
def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        return json.load(f)


def test_logq_fixture_is_the_reference_literal(fixtures):
    lits = _literals(os.path.join(REF, "test_layers.py"))
    g = fixtures["logq"]
    assert g["logits"] in lits and g["expected"] in lits and g["candidate_prob_lookup"] in lits
    assert [g["candidate_ids"]] in lits                                   # tf.constant([["id1", "id2", "id3"]], shape=(3, 1))


def test_brute_force_fixture_is_the_reference_literal(fixtures):
    lits = _literals(os.path.join(REF, "test_indices.py"))
    g = fixtures["brute_force"]
    assert g["query_vocab"] in lits and g["query_table"] in lits and g["candidate_ids"] in lits
    assert g["queries"] in lits and g["expected"] in lits
    for row in g["candidate_embeddings"]:                                  # one tf.constant([x, y]) per candidate
        assert row in lits
    with open(os.path.join(REF, "test_indices.py")) as f:
        src = f.read()
    assert f"BruteForceIndex({g['k']}," in src and f"ds.batch({g['candidate_batch']})" in src


def test_recall_fixture_is_the_reference_literal(fixtures):
    path = os.path.join(REF, "test_recall.py")
    lits = _literals(path)
    flat = [[x.decode() if isinstance(x, bytes) else x for x in _flat(v)] for v in lits if isinstance(v, (list, tuple))]   # b"id1" literals
    g = fixtures["recall"]
    for key in ("query_ids", "true_candidate_ids", "static_candidates"):
        assert g[key] in flat, key
    with open(path) as f:
        src = f.read()
    assert g["ks"] in lits and f"ks={g['ks']}" in src
    for k, v in g["expected"].items():
        assert re.search(rf"\b{k}:\s*tf\.constant\({re.escape(str(v))},", src), (k, v)      # 1: tf.constant(0.2, dtype=tf.float64)
