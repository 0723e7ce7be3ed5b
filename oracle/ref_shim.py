"""Import shim for the LIVE reference (/root/reference) — test infrastructure only.

The reference (modelscope/FunASR, pure Python/PyTorch) imports four pip packages
that are absent from this image but never *called* on the offline Paraformer path
(SURVEY.md §8c / App. B): rapidfuzz, kaldiio, librosa, omegaconf.  Empty stub
modules make `import funasr` succeed so the reference's own classes can be run on
CPU to (a) validate oracle/paraformer_oracle.py and (b) generate tests/golden/*.

/root/reference does not exist on the GPU box.  What CAN travel there is the offline install of the unmodified
reference under the git-ignored `baseline/_ref/` (`pip install --no-index --no-deps --target baseline/_ref /root/reference`,
DESIGN.md §8): bench.py's `--impl reference` arm (oracle/ref_runner.py) and the one GPU test that drives
`AutoModel.generate()` through this backend import it from there.  Nothing else under tests -m gpu, bench.py's own arm or
__graft_entry__.smoke() touches the reference.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    cands = [os.environ.get("FUNASR_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "funasr")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "funasr"))


def reference_kind() -> str:
    """'tree' = the read-only source tree (/root/reference), 'installed' = the pip --target copy under baseline/_ref."""
    return "installed" if os.path.abspath(REFERENCE_ROOT).startswith(os.path.join(_REPO, "baseline")) else "tree"


def install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:  # OmegaConf placeholder: only attribute access happens at import
        @staticmethod
        def create(x=None):
            return x if x is not None else {}

        @staticmethod
        def to_container(x, **kw):
            return x

    rf = stub("rapidfuzz")
    rf.distance = stub("rapidfuzz.distance", Levenshtein=object())
    rf.fuzz = stub("rapidfuzz.fuzz")
    stub("kaldiio")
    stub("librosa")
    stub("omegaconf", DictConfig=dict, ListConfig=list, OmegaConf=_Dummy)


def import_reference():
    """Returns the imported reference `funasr` package (raises if unavailable)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging
    logging.disable(logging.WARNING)
    import funasr  # noqa: F401
    logging.disable(logging.NOTSET)
    return funasr
