"""Import the unmodified reference modules: from /root/reference in the build container, otherwise from the
installed copy baseline/_ref/ (baseline/install_reference.py: a verbatim copy, git-ignored, that travels to the GPU box).

TEST INFRASTRUCTURE.  Used by `oracle/gen_golden.py` (to produce tests/golden/*.npz), by the CPU-only tests that are
skipped when no tree is present, by the config-4 GPU test (the drop-in inside the reference VQVAE) and by
`bench.py --impl reference` / the `cpu_baseline` leg (the reference's own modules timed on the host cores).

src/util.py:7-12 imports editdistance, soundfile and matplotlib, none of which is
installed; empty stub modules are enough because the quantizer never calls them.
"""
import contextlib
import os
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _pick_root():
    env = os.environ.get("VQB_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _INSTALLED):
        if os.path.isfile(os.path.join(cand, "src", "embed.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()

_STUBS = ["editdistance", "soundfile", "matplotlib", "matplotlib.pyplot",
          "tensorboardX", "librosa"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "embed.py"))


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    tbx = sys.modules["tensorboardX"]
    if not hasattr(tbx, "SummaryWriter"):
        tbx.SummaryWriter = object


@contextlib.contextmanager
def reference_cwd():
    """The YAML's `phn_attr_pth: 'data/phn_attr.csv'` is relative to the reference root."""
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def import_reference():
    """Returns the reference `src.embed` module (L2Embedding, SeperateEmbedding, neg_batch_l2)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True          # the reference tree is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.embed as ref_embed           # noqa: E402
    return ref_embed


def import_reference_vqvae():
    import_reference()
    import src.vqvae as ref_vqvae           # noqa: E402
    return ref_vqvae


def load_codebook_cfg(yaml_name: str) -> dict:
    """`model.codebook` block of a reference YAML, as VQVAE.__init__ sees it (src/vqvae.py:41)."""
    import copy
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "config", yaml_name)) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    return copy.deepcopy(cfg["model"]["codebook"])
