"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for the reference Trainer's host-side batch chunking
(distributed_trainer.py:77-169 `calculate_chunk_sizes`, `split_dict_lists`; :221-230 `merge_candidates`), produced by
calling the REFERENCE's own static methods (third-party imports stubbed as in make_golden.py).  Runs only in the build
container; writes tests/golden/trainer_chunks.json, which is committed.

Usage: python oracle/make_golden_trainer.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import REF, install_stubs  # noqa: E402


def main():
    install_stubs()
    import types
    sys.modules.setdefault("tqdm", types.ModuleType("tqdm")).tqdm = lambda x, **k: x
    tr = sys.modules.setdefault("transformers", types.ModuleType("transformers"))
    if not hasattr(tr, "GenerationConfig"):
        tr.GenerationConfig = lambda **k: types.SimpleNamespace(**k)
    sys.path.insert(0, REF)
    import distributed_trainer as dt
    cases = []
    grid = [(30, 2, 1, 8), (512, 4, 4, 8), (8, 0, 1, 8), (7, 2, 1, 8), (3, 4, 2, 8), (2, 4, 1, 1), (16, 3, 2, 4), (64, 4, 4, 0),
            (5, 2, 3, 1), (1, 1, 1, 1), (9, 2, 2, 4), (10, 0, 2, 5), (4, 4, 1, 8), (33, 5, 1, 8), (12, 1, 4, 2)]
    for (bs, na, nl, lc) in grid:
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                out = dt.Trainer.calculate_chunk_sizes(bs, na, nl, lc)
                err = None
            except Exception as e:  # noqa: BLE001
                out, err = None, type(e).__name__
        cases.append({"batch_size": bs, "num_actors": na, "num_learners": nl, "learner_chunk_size": lc, "chunks": out, "error": err})
    data = {"problem": [f"p{i}" for i in range(10)], "solution": [f"s{i}" for i in range(10)]}
    splits = []
    for sizes in ([4, 3, 3], [10], 10, [1, 9], [5, 4]):
        try:
            out = dt.Trainer.split_dict_lists(data, sizes)
            err = None
        except Exception as e:  # noqa: BLE001
            out, err = None, type(e).__name__
        splits.append({"sizes": sizes, "out": out, "error": err})
    # the CLI surface: every flag of the reference's train_distributed.py with its type and default, taken from the source
    # text (the module itself cannot be imported: it downloads a dataset at import time)
    import ast
    flags = []
    tree = ast.parse(open(os.path.join(REF, "train_distributed.py")).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            kw = {k.arg: k.value for k in node.keywords}
            flags.append({"flag": node.args[0].value,
                          "type": kw["type"].id if "type" in kw else None,
                          "default": ast.literal_eval(kw["default"]) if "default" in kw else None,
                          "choices": ast.literal_eval(kw["choices"]) if "choices" in kw else None})
    path = os.path.join(ROOT, "tests", "golden", "trainer_chunks.json")
    json.dump({"source": "BY571/DistRL-LLM distributed_trainer.py Trainer.calculate_chunk_sizes / split_dict_lists (run verbatim)",
               "chunk_cases": cases, "split_data": data, "split_cases": splits, "cli_flags": flags}, open(path, "w"), indent=1)
    print("wrote", path, len(cases), "chunk cases")


if __name__ == "__main__":
    main()
