"""Parity numbers the GPU tests measure, kept on file: every call merges one entry into a JSON document
(default gpurun_out/r2_parity.json — gpurun brings that directory back; the committed copy is
profiles/r2_parity.json). The tests assert on the same numbers; this is the record of what they saw."""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.environ.get("RTX_PARITY_JSON") or os.path.join(ROOT, "gpurun_out", "r2_parity.json")


def record(section: str, key: str, value) -> None:
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        doc = {}
        if os.path.exists(PATH):
            with open(PATH) as f:
                doc = json.load(f)
        doc.setdefault(section, {})[str(key)] = value
        with open(PATH + ".tmp", "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True)
        os.replace(PATH + ".tmp", PATH)
    except OSError:
        pass  # a read-only checkout: the assertions still ran
