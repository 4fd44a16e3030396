"""Install the UNMODIFIED reference (zmoon/crt1d) into the git-ignored `baseline/_ref/` so that `bench.py --impl
reference` can time the reference's own solver functions on the GPU box (where /root/reference does not exist;
`baseline/_ref` travels with the repo snapshot).  Called by `__graft_entry__.build()` in the build container.

Recipe (the base contract's offline install): copy the read-only tree to /tmp (the build writes `crt1d/_version.py`
and egg-info into the source tree), then
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>
`--no-deps`: xarray and matplotlib are not in this image (the solver modules do not need them).  setuptools_scm is
absent too, so the wheel carries only the .py files; the package DATA the reference's MANIFEST would add
(`variables.yml`, `data/*.txt|csv`) is copied afterwards, byte for byte.  Nothing of this is tracked by git.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("CRT1D_REFERENCE", "/root/reference")


def staged():
    return os.path.isfile(os.path.join(REF_DST, "crt1d", "solvers", "_solve_2s.py")) and \
        os.path.isfile(os.path.join(REF_DST, "crt1d", "variables.yml"))


def stage(force=False, verbose=True):
    """Returns a status string; never raises (the reference arm falls back to the oracle port if this fails)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "crt1d", "solvers")):
        return "reference tree absent (GPU box): using the prebuilt baseline/_ref" if staged() else "reference tree absent, baseline/_ref missing"
    if staged() and not force:
        return "baseline/_ref already staged"
    tmp = tempfile.mkdtemp(prefix="crt1d_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE_ROOT, src)
        shutil.rmtree(REF_DST, ignore_errors=True)
        env = dict(os.environ, SETUPTOOLS_SCM_PRETEND_VERSION="0.0.0")
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_DST, src]
        proc = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tmp)
        how = "pip install --no-deps --target"
        if proc.returncode != 0 or not os.path.isdir(os.path.join(REF_DST, "crt1d")):
            # last resort: the package directory as it lies in the reference tree
            how = "copy of the package directory (pip failed: %s)" % proc.stderr.strip().splitlines()[-1:] 
            shutil.rmtree(REF_DST, ignore_errors=True)
            shutil.copytree(os.path.join(REFERENCE_ROOT, "crt1d"), os.path.join(REF_DST, "crt1d"))
        n_data = 0
        for root, _, files in os.walk(os.path.join(REFERENCE_ROOT, "crt1d")):
            rel = os.path.relpath(root, REFERENCE_ROOT)
            for f in files:
                if f.endswith((".py", ".pyc")):
                    continue
                dst = os.path.join(REF_DST, rel, f)
                if not os.path.exists(dst):
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), dst)
                    n_data += 1
        with open(os.path.join(REF_DST, "STAGED.json"), "w") as fh:
            json.dump({"how": how, "package_data_files_added": n_data, "source": REFERENCE_ROOT}, fh)
        msg = f"staged the reference into baseline/_ref ({how}; {n_data} package-data files added)"
    except Exception as e:  # noqa: BLE001
        msg = f"staging failed: {e!r}"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print(f"[build] {msg}")
    return msg


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv, verbose=False))
