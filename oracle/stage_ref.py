"""Stages the reference's own DenseNet module for the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/stage_ref.py          (run by __graft_entry__.build() whenever /root/reference is present)

`/root/reference` does not exist on the GPU box, but `bench.py --impl reference` should time the REFERENCE's code, not a port, where
that is possible.  RegressionNetwork/DenseNet.py imports unchanged (torch only); this recipe places an UNMODIFIED copy in
`oracle/_ref/` -- git-ignored (never committed: no reference source enters the history), not gpurun-ignored (it travels with the
snapshot like a built .so).  The SHA-256 of the staged file is recorded next to it so that the bench line can say exactly what it ran.
Nothing in the product imports it; `bench.py`'s reference arm / cpu_baseline leg uses it when present and falls back to the port
(oracle/densenet_oracle.py, the same ATen ops in the same order) otherwise.  The render half of the path stays the port: the
reference's RegressionNetwork/util.py is not importable (merge-conflict markers at :5-15, :248-285)."""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference/RegressionNetwork/DenseNet.py"
DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def stage():
    if not os.path.exists(SRC):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    dst = os.path.join(DST_DIR, "DenseNet.py")
    shutil.copyfile(SRC, dst)
    h = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "DenseNet.sha256"), "w") as f:
        f.write(h + "  RegressionNetwork/DenseNet.py (unmodified copy staged by oracle/stage_ref.py)\n")
    return dst


def load():
    """The staged reference module (or None): imported under a private name so that it can never shadow the product's `DenseNet`."""
    path = os.path.join(DST_DIR, "DenseNet.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_emlight_reference_DenseNet", path)
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(stage())
