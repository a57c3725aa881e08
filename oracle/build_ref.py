"""Recipe for ``oracle/_ref``: the UNMODIFIED reference, installed so that it travels to the GPU box.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py [--force]

The reference (broadinstitute/permutect, /root/reference) is pure Python, so "building" it is a plain
``pip install --no-deps --target oracle/_ref`` from a scratch copy of the checkout (the checkout itself is read-only
and setuptools wants to write an egg-info next to the sources).  No reference source is copied into the tracked
tree: ``oracle/_ref/`` is git-ignored (it is NOT gpurun-ignored, so the GPU box receives it like the built ``.so``).
Three of its import-time dependencies are absent from this image and unused on the ArtifactModel path
(cyvcf2, intervaltree, matplotlib; pymc for the posterior M step) — ``oracle/ref_stubs`` satisfies those imports.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may use what this installs
(through ``oracle/reference.py``); nothing under ``permutect_b200/`` does.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("PERMUTECT_REFERENCE", "/root/reference")
STAMP = os.path.join(TARGET, ".installed_from")


def available() -> bool:
    return os.path.isdir(os.path.join(TARGET, "permutect"))


def build(force: bool = False) -> str:
    """Installs the reference into oracle/_ref.  Returns the target path; raises if the checkout is absent and nothing
    was installed before (on the GPU box the prebuilt directory is used as it arrived)."""
    if available() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "permutect")):
        raise RuntimeError(f"reference checkout not found at {SOURCE} and oracle/_ref is empty")
    scratch = tempfile.mkdtemp(prefix="permutect_ref_")
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "integration-tests", "*.pt", "*.tar"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not available():
            # same files, without pip: the package is pure Python
            os.makedirs(TARGET, exist_ok=True)
            shutil.copytree(os.path.join(SOURCE, "permutect"), os.path.join(TARGET, "permutect"), dirs_exist_ok=True)
        with open(STAMP, "w") as f:
            f.write(SOURCE + "\n")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
