"""Drop-in boundary: every module, class, method and function of the reference's ``pkg`` that is in scope exists here
under the same name with the same leading argument names.  Compared by parsing both trees with ``ast`` (the reference imports
TensorFlow and cannot be imported); runs where the reference checkout is present."""
import ast
import os

import pytest

REF = "/root/reference/pkg"
OURS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hm-retrieval-two-tower_b200", "pkg")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

# the only symbols without a namesake: two private helpers folded into their callers
OUT_OF_SCOPE_FILES = set()
OUT_OF_SCOPE_NAMES = {("schema/features.py", "Feature._init_vocab"),
                      ("modelling/tfrecord_dataset.py", "TFRecordDatasetFactory._parse_function")}


def _signatures(path, follow=True):
    with open(path) as f:
        tree = ast.parse(f.read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            out[node.name] = None
            fields = [st.target.id for st in node.body if isinstance(st, ast.AnnAssign) and isinstance(st.target, ast.Name)]
            if fields:                                  # dataclass fields: their order is the positional constructor signature
                out[f"{node.name}.<fields>"] = fields
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef):
                    out[f"{node.name}.{fn.name}"] = [a.arg for a in fn.args.args]
        elif isinstance(node, ast.FunctionDef):
            out[node.name] = [a.arg for a in node.args.args]
        elif isinstance(node, ast.ImportFrom) and follow and node.module and node.module.startswith("pkg."):
            # a module that re-exports a class defined next door (schema/model_config.py -> schema/config.py)
            target = os.path.join(os.path.dirname(OURS), *node.module.split(".")) + ".py"
            if not os.path.exists(target):        # ``from pkg.schema import dtypes``: a package, nothing to follow
                continue
            src = _signatures(target, follow=False)
            for alias in node.names:
                for sym, args in src.items():
                    if sym == alias.name or sym.startswith(alias.name + "."):
                        out.setdefault(sym, args)
    return out


def test_every_in_scope_reference_symbol_exists_with_the_same_arguments():
    problems, checked = [], 0
    for root, _, files in os.walk(REF):
        for name in files:
            if not name.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, name), REF)
            if rel in OUT_OF_SCOPE_FILES:
                continue
            mine = os.path.join(OURS, rel)
            if not os.path.exists(mine):
                problems.append(f"missing module {rel}")
                continue
            ref_sigs, our_sigs = _signatures(os.path.join(REF, rel), follow=False), _signatures(mine)
            for sym, args in ref_sigs.items():
                if (rel, sym) in OUT_OF_SCOPE_NAMES:
                    continue
                checked += 1
                if sym not in our_sigs:
                    problems.append(f"{rel}: missing {sym}")
                elif args is not None and our_sigs[sym][:len(args)] != args:      # extra trailing (defaulted) arguments are allowed
                    problems.append(f"{rel}: {sym}{args} here is {our_sigs[sym]}")
    assert not problems, "\n".join(problems)
    assert checked > 60
