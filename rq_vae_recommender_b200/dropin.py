"""Make the UNMODIFIED reference scripts run on the B200 modules.

    import rq_vae_recommender_b200.dropin as dropin
    dropin.install(reference_root="/path/to/RQ-VAE-Recommender")   # before `import train_rqvae`
    import train_rqvae; train_rqvae.train(...)

``install`` pre-seeds ``sys.modules`` so that every ``from modules.quantize import ...`` / ``from init.kmeans import
...`` inside the reference (train_rqvae.py:13-15, modules/tokenizer/semids.py:10, train_decoder.py:13-20) resolves to
the replacement modules of this package; everything else (data/, modules/model.py, evaluate/, the scripts) is imported
from the reference tree untouched.  Checkpoints pickle ``modules.quantize.Quantize`` etc. by module path,
so ``torch.load(..., weights_only=False)`` of the shipped files also lands on the replacement classes.
gin-config is not in this image: a small compatible shim is registered as ``gin`` when the real one is missing.
"""
import importlib
import sys
import types

_ALIASES = {
    "modules.quantize": "rq_vae_recommender_b200.modules.quantize",
    "modules.rqvae": "rq_vae_recommender_b200.modules.rqvae",
    "modules.encoder": "rq_vae_recommender_b200.modules.encoder",
    "modules.loss": "rq_vae_recommender_b200.modules.loss",
    "modules.normalize": "rq_vae_recommender_b200.modules.normalize",
    "init.kmeans": "rq_vae_recommender_b200.init.kmeans",
    "distributions.gumbel": "rq_vae_recommender_b200.distributions.gumbel",
}
_TOKENIZER = ("modules.tokenizer.semids", "rq_vae_recommender_b200.modules.tokenizer.semids")


def install(reference_root=None, replace_tokenizer=True, gin_shim=True):
    if gin_shim and "gin" not in sys.modules:
        try:
            import gin  # noqa: F401
        except ImportError:
            from . import gin_compat
            sys.modules["gin"] = gin_compat
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    for parent in ("init", "distributions"):          # namespace packages in the reference (no __init__.py)
        if parent not in sys.modules and reference_root is None:
            sys.modules[parent] = types.ModuleType(parent)
            sys.modules[parent].__path__ = []
    for alias, real in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(real)
    if replace_tokenizer:
        sys.modules[_TOKENIZER[0]] = importlib.import_module(_TOKENIZER[1])
    return sorted(list(_ALIASES) + ([_TOKENIZER[0]] if replace_tokenizer else []))


def uninstall():
    for alias in list(_ALIASES) + [_TOKENIZER[0]]:
        mod = sys.modules.get(alias)
        if mod is not None and mod.__name__.startswith("rq_vae_recommender_b200"):
            del sys.modules[alias]
