"""Recipe for `oracle/_ref/`: the UNMODIFIED reference package, installed from `/root/reference` by pip.

TEST / BENCH INFRASTRUCTURE ONLY.  `oracle/_ref/` is git-ignored (never in history) but NOT gpurun-ignored,
so the installed copy travels to the GPU box, where `/root/reference` does not exist.  It is used
  * by `bench.py --impl reference` and the `cpu_baseline` leg: the reference's own `create_model` /
    `create_anchors` / `Decoder` / `extract_targets` / `MultiBoxLoss` timed on the host cores
    (`cpu_baseline.kind == "reference"`), through `oracle/ref_runner.py`;
  * by tests that pin the oracle restatement against the real thing when it is present.
The product (`ssds_pytorch_b200/`) never imports it.

    python oracle/build_ref.py          # no-op when oracle/_ref is already there or /root/reference is absent

The reference's setup.py writes build/ and egg-info into its source tree and `/root/reference` is read-only,
so the install runs from a scratch copy under /tmp; `--no-deps` because its requirements (DALI, apex, cv2 pins)
are not resolvable offline and none is on the timed path.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("SSDS_REFERENCE", "/root/reference")


def have_ref():
    return os.path.isfile(os.path.join(TARGET, "ssds", "modeling", "model_builder.py"))


def build(force=False, verbose=False):
    """Returns the target dir if the reference is installed (now or before), else None."""
    if have_ref() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE, "ssds")):
        return TARGET if have_ref() else None
    scratch = tempfile.mkdtemp(prefix="ssds_ref_src_")
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "doc", "*.jpg", "*.png"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed")
        # the shipped test config (experiments/cfgs/tests/test.yml) is data the plumbing config reads
        yml = os.path.join(REFERENCE, "experiments", "cfgs", "tests", "test.yml")
        if os.path.isfile(yml):
            shutil.copy(yml, os.path.join(TARGET, "test.yml"))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
