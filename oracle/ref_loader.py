"""Import the UNMODIFIED reference (read-only, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY. Used by oracle/make_goldens.py and by CPU tests that
run where /root/reference exists, to pin oracle/port.py against the reference's
own code (SURVEY.md §8c recipe).  Never imported by the product package.

The reference drags in packages that are absent from this image (omegaconf,
hydra, nltk, timm, albumentations); none of them touch the arithmetic of the
hot path, so they are replaced by empty stub modules before the import.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MCLIP_REFERENCE_ROOT", "/root/reference")
REF_CODEBASE = os.path.join(REF_ROOT, "src", "codebase")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODEBASE, "breastclip"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _DictConfig(dict):
        pass

    class _OmegaConf:
        @staticmethod
        def to_yaml(cfg):
            return str(cfg)

        @staticmethod
        def resolve(cfg):
            return None

        @staticmethod
        def to_container(cfg, **kw):
            return cfg

    _stub("omegaconf", DictConfig=_DictConfig, OmegaConf=_OmegaConf)

    def _hydra_main(*a, **k):
        def deco(fn):
            return fn
        return deco

    _stub("hydra", main=_hydra_main)
    nltk = _stub("nltk", download=lambda *a, **k: True)
    tok = _stub("nltk.tokenize", sent_tokenize=lambda s: [s], RegexpTokenizer=object)
    nltk.tokenize = tok

    def _no_timm(*a, **k):
        raise RuntimeError("timm is not installed in this image")

    _stub("timm", create_model=_no_timm)
    alb = _stub("albumentations", __all__=[])
    alb.pytorch = _stub("albumentations.pytorch", ToTensorV2=object)


_REF = None


def load_reference():
    """Returns the imported `breastclip` package of the reference."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError(f"reference not present at {REF_CODEBASE}")
    import transformers  # noqa: F401  must be imported BEFORE the timm stub exists (it probes find_spec("timm"))
    from transformers import AutoTokenizer, AutoModel, AutoConfig, BertModel  # noqa: F401  resolve lazies now
    _install_stubs()
    if REF_CODEBASE not in sys.path:
        sys.path.insert(0, REF_CODEBASE)
    import breastclip  # noqa: E402  (the reference package)
    from breastclip.model.modules import efficientnet_custom as _enc
    # EfficientNet.from_pretrained downloads ImageNet weights (utils:597-602); there is no
    # network, so weight loading is neutralised (random init is seeded by the caller).
    _enc.load_pretrained_weights = lambda *a, **k: None

    class _NullWriter:
        def add_scalar(self, *a, **k):
            pass

    env = breastclip.util.GlobalEnv.get()
    env.summary_writer.train = _NullWriter()
    env.summary_writer.valid = _NullWriter()
    _REF = breastclip
    return breastclip


def reset_global_env():
    """GlobalEnv is a singleton built at first use; after init_process_group it must be rebuilt."""
    bc = load_reference()
    bc.util.GlobalEnv._instance = None
    env = bc.util.GlobalEnv.get()

    class _NullWriter:
        def add_scalar(self, *a, **k):
            pass

    env.summary_writer.train = _NullWriter()
    env.summary_writer.valid = _NullWriter()
    return env
