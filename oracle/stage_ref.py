"""Stage the UNMODIFIED reference package next to the oracle.   *** TEST / BENCH INFRASTRUCTURE ***

The reference is pure Python: there is nothing to compile, and its own build backend (hatchling)
and its Lightning / Hydra dependencies are absent from the image, so ``pip install --target``
cannot run.  This recipe does what that install would do for a pure-Python wheel: it copies the
package's ``.py`` files, byte for byte, from ``/root/reference/src/generative_recommenders_pl`` into
the git-ignored ``oracle/_ref/generative_recommenders_pl`` and writes a manifest of their SHA-256
sums.  ``oracle/_ref`` is NOT gpurun-ignored, so the staged copy travels to the GPU box, where
``bench.py --impl reference`` / ``cpu_baseline`` import the reference's own modules from it
(``oracle/ref_verbatim.py``; kind "reference").  Nothing under ``mygenerativerecommenders_b200/``
reads it, and nothing of it enters the git history.

    python -m oracle.stage_ref            # run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

SRC = Path("/root/reference/src/generative_recommenders_pl")
DST = Path(__file__).resolve().parent / "_ref"


def stage() -> Path | None:
    if not SRC.is_dir():
        return DST if (DST / "MANIFEST.json").exists() else None
    pkg = DST / "generative_recommenders_pl"
    if pkg.exists():
        shutil.rmtree(pkg)
    manifest = {}
    for f in sorted(SRC.rglob("*.py")):
        rel = f.relative_to(SRC)
        out = pkg / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, out)
        manifest[str(rel)] = hashlib.sha256(f.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps(
        {"source": str(SRC), "files": manifest}, indent=1, sort_keys=True))
    return DST


if __name__ == "__main__":
    print(stage())
