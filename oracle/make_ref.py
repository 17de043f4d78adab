"""Stage the UNMODIFIED reference under oracle/_ref/ so it can travel to the GPU box (TEST / BENCH INFRASTRUCTURE ONLY).

/root/reference exists in the build container only.  The GPU-side drop-in tests (the reference's own trainer.train() driving this
library's modules), the `--impl reference` arm of bench.py and the "reference through PyTorch/cuDNN on the B200" extra need the
reference itself, not a restatement.  This recipe copies the Python sources of the path and of its callers, byte for byte, from where
they lie into oracle/_ref/ (git-ignored: never committed; not gpurun-ignored: it travels with the snapshot like a built .so).
__graft_entry__.build() runs it whenever /root/reference is present.  Nothing in the product package imports oracle/ or oracle/_ref
(tests/test_host_cpu.py::test_product_does_not_import_oracle).

  python oracle/make_ref.py            # (re)stage;  prints the file list with sha256 prefixes
"""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = (
    "main.py", "trainer.py", "utils.py", "metrics.py", "GPUtilext.py",
    "model_zoo/nconv.py", "model_zoo/nconv_sd.py", "model_zoo/confidence_depth_framework.py",
    "model_zoo/s2d_u_resnet.py", "model_zoo/s2d_resnet.py",
    "dataloaders/dense_to_sparse.py", "dataloaders/dataloader_factory.py", "dataloaders/transforms.py",
)


def stage(verbose=False):
    """Copies FILES from SRC to DST when SRC exists; returns the list of staged paths (empty when there is no reference here)."""
    if not os.path.isdir(os.path.join(SRC, "model_zoo")):
        return []
    out = []
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        data = open(s, "rb").read()
        if not (os.path.exists(d) and open(d, "rb").read() == data):
            if os.path.exists(d):
                os.chmod(d, 0o644)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
        out.append(d)
        if verbose:
            print("%s  %s" % (hashlib.sha256(data).hexdigest()[:16], rel))
    return out


if __name__ == "__main__":
    n = stage(verbose=True)
    print("staged %d file(s) into %s" % (len(n), DST) if n else "no reference at %s: nothing staged" % SRC)
    sys.exit(0)
