"""TEST INFRASTRUCTURE ONLY — run the reference's OWN, unmodified Python files (model.py, batch.py, util.py).

The reference is pure Python (no C/C++ to compile), so "building" `oracle/_ref` means staging the few files of the
path where the GPU box can see them: `/root/reference` exists in the build container only.

    python oracle/reference_runner.py stage      # /root/reference/{chem,bio}/<files> -> oracle/_ref/{chem,bio}/

`oracle/_ref/` is git-ignored (never part of the history: no reference source is committed) but not gpurun-ignored,
so it travels with the snapshot like a built `.so`.  `__graft_entry__.build()` runs the staging step whenever
`/root/reference` is present.  Nothing under `pretrain-gnns_b200/` may import this module; users are `tests/`,
`tests/golden/make_golden*.py` and `bench.py`'s `--impl reference` / `cpu_baseline` legs (kind "reference").

The third-party modules holding the reference's arithmetic (torch_geometric 1.0.3, torch_scatter 1.1.2; also rdkit and
tensorboardX at import time) are not installable here; `oracle/pyg103_standin` serves the handful of names the files
import (its README states the semantics it encodes and which of them are recalled, not verifiable).
"""
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_STAGED = os.path.join(HERE, "_ref")
STANDIN = os.path.join(HERE, "pyg103_standin")
FILES = ("model.py", "batch.py", "loader.py", "dataloader.py", "util.py")
# shipped checkpoints behind tests/golden/pretrained.npz (tests/golden_util.PRETRAINED): staged, never committed
WEIGHTS = ("chem/model_gin/masking.pth", "chem/model_architecture/gcn_contextpred.pth", "chem/model_architecture/gat_contextpred.pth",
           "chem/model_architecture/graphsage_contextpred.pth", "bio/model_gin/masking.pth")
_LOCAL_MODULES = ("model", "loader", "dataloader", "batch", "util", "splitters")


def stage(verbose=True):
    """Copy the path's reference files into oracle/_ref (git-ignored).  No-op when /root/reference is absent."""
    if not os.path.isdir(REF_SRC):
        return False
    for domain in ("chem", "bio"):
        dst = os.path.join(REF_STAGED, domain)
        os.makedirs(dst, exist_ok=True)
        for f in FILES:
            src = os.path.join(REF_SRC, domain, f)
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(dst, f))
    for f in WEIGHTS:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_STAGED, "weights", f)
        if os.path.exists(src) and not (os.path.exists(dst) and os.path.getsize(dst) == os.path.getsize(src)):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    with open(os.path.join(REF_STAGED, "PROVENANCE.txt"), "w") as fh:
        fh.write("byte copies of /root/reference/{chem,bio}/{%s} and, under weights/, of the checkpoints %s; staged by "
                 "oracle/reference_runner.py; git-ignored\n" % (",".join(FILES), ", ".join(WEIGHTS)))
    if verbose:
        print("staged reference files under", REF_STAGED)
    return True


def root():
    """Directory holding {chem,bio}/model.py of the reference: the original tree if present, else the staged copy."""
    if os.path.isdir(os.path.join(REF_SRC, "chem")):
        return REF_SRC
    if os.path.isfile(os.path.join(REF_STAGED, "chem", "model.py")):
        return REF_STAGED
    return None


def available():
    return root() is not None


class _Placeholder:
    """Stands for any name of an absent package: attribute access yields another placeholder (loader.py builds tables of
    `Chem.rdchem.ChiralType.*` constants at import time), calling one raises."""

    def __init__(self, path):
        self._path = path

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder(self._path + "." + name)

    def __call__(self, *a, **k):
        raise RuntimeError("%s is a stub: the real package is not installed in this image" % self._path)

    def __repr__(self):
        return "<stub %s>" % self._path


class _StubModule(types.ModuleType):
    """`from rdkit.Chem.rdMolDescriptors import GetMorganFingerprintAsBitVect` succeeds at import time."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder(self.__name__ + "." + name)


def _stub_missing():
    """rdkit / tensorboardX are imported at module top by loader.py / the scripts but never reached on this path."""
    for name in ("rdkit", "rdkit.Chem", "rdkit.Chem.Descriptors", "rdkit.Chem.AllChem", "rdkit.DataStructs",
                 "rdkit.Chem.rdMolDescriptors", "rdkit.Chem.Scaffolds", "rdkit.Chem.Scaffolds.MurckoScaffold", "tensorboardX"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = _StubModule(name)
            m.__stub__ = True
            sys.modules[name] = m
            if "." in name:
                parent, _, leaf = name.rpartition(".")
                setattr(sys.modules[parent], leaf, m)
    rd = sys.modules.get("rdkit")
    if rd is not None and getattr(rd, "__stub__", False):
        for leaf in ("Chem", "DataStructs"):
            setattr(rd, leaf, sys.modules["rdkit." + leaf])


def load(domain, module="model"):
    """Import `<root>/<domain>/<module>.py` (and whatever sibling files it imports) with the stand-in on sys.path.
    Returns the module object; the sibling modules are dropped from sys.modules afterwards so chem and bio can both
    be loaded in one process."""
    r = root()
    if r is None:
        raise RuntimeError("the reference is not available: neither /root/reference nor oracle/_ref exists "
                           "(run `python oracle/reference_runner.py stage` in the build container)")
    if STANDIN not in sys.path:
        sys.path.insert(0, STANDIN)
    _stub_missing()
    for m in _LOCAL_MODULES:
        sys.modules.pop(m, None)
    sys.path.insert(0, os.path.join(r, domain))
    try:
        mod = importlib.import_module(module)
    finally:
        sys.path.pop(0)
        for m in _LOCAL_MODULES:
            sys.modules.pop(m, None)
    return mod


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stage":
        sys.exit(0 if stage() else 1)
    print("root:", root())
