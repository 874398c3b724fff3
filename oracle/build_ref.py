"""ORACLE - TEST INFRASTRUCTURE ONLY.  Vendors the UNMODIFIED reference into ``oracle/_ref/``.

    python oracle/build_ref.py            # copy + verify (needs /root/reference, i.e. the build container)
    python oracle/build_ref.py --check    # verify an existing oracle/_ref against the committed manifest

``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the
tree travels to the GPU box like the built ``.so`` files.  ``bench.py --impl reference`` runs the reference's own
``GraphCreator`` from there over the import shims in ``oracle/shims`` - the reference CPU path, timed on the
box's host cores.  The committed ``oracle/ref_manifest.json`` holds the sha256 of every file of
``/root/reference/src/anemoi`` at the time the manifest was written; the copy is byte-for-byte (``shutil.copy2``)
and is re-verified against the manifest before every reference-arm run, so "unmodified" is checkable.
"""

from __future__ import annotations

import hashlib
import json
import pathlib
import shutil
import sys

ORACLE = pathlib.Path(__file__).resolve().parent
SRC = pathlib.Path("/root/reference/src/anemoi")
DST = ORACLE / "_ref" / "anemoi"
MANIFEST = ORACLE / "ref_manifest.json"


def _sha(path: pathlib.Path) -> str:
    return hashlib.sha256(path.read_bytes()).hexdigest()


def tree_manifest(root: pathlib.Path) -> dict[str, str]:
    files = sorted(p for p in root.rglob("*") if p.is_file() and "__pycache__" not in p.parts)
    return {str(p.relative_to(root)): _sha(p) for p in files}


def check(root: pathlib.Path = DST) -> dict[str, str]:
    """Raise unless ``root`` holds exactly the files of the committed manifest, byte for byte."""
    want = json.loads(MANIFEST.read_text())["files"]
    have = tree_manifest(root)
    if have != want:
        missing = sorted(set(want) - set(have))
        extra = sorted(set(have) - set(want))
        changed = sorted(k for k in set(want) & set(have) if want[k] != have[k])
        raise RuntimeError(f"oracle/_ref is not the unmodified reference: missing {missing}, extra {extra}, changed {changed}")
    return have


def build(write_manifest: bool = False) -> pathlib.Path:
    """Copy the reference package; with ``write_manifest`` also (re)write ``ref_manifest.json`` from the source."""
    if not SRC.exists():
        if DST.exists():
            check()
            return DST
        raise RuntimeError(f"{SRC} not found and {DST} not built: run this in the build container")
    if write_manifest or not MANIFEST.exists():
        MANIFEST.write_text(json.dumps({"source": str(SRC), "files": tree_manifest(SRC)}, indent=1, sort_keys=True) + "\n")
    if DST.exists():
        shutil.rmtree(DST)
    DST.parent.mkdir(exist_ok=True)
    shutil.copytree(SRC, DST, copy_function=shutil.copy2, ignore=shutil.ignore_patterns("__pycache__"))
    check()
    return DST


if __name__ == "__main__":
    if "--check" in sys.argv:
        print(f"{len(check())} files match {MANIFEST.name}")
    else:
        print(build(write_manifest="--write-manifest" in sys.argv))
