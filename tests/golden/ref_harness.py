"""Loader for the UNMODIFIED reference (only usable where /root/reference exists, i.e. the build container).

Installs in-memory stand-ins for the packages the reference imports but this image lacks
(gin, accelerate, torch_geometric, polars, sentence_transformers  --  SURVEY appendix C), puts
/root/reference on sys.path and disables torch.compile (RqVae.forward is decorated, rqvae.py:141;
eager fp32 CPU is the canonical oracle).  Nothing here is shipped or used on the GPU box.
"""
import os
import sys
import types

REFERENCE = os.environ.get("RQ_REFERENCE_PATH", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "modules"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")
    if "gin" not in sys.modules:
        ident = lambda x=None, *a, **k: x if x is not None else (lambda f: f)
        _stub("gin", configurable=ident, constants_from_enum=ident,
              parse_config_file=lambda *a, **k: None)
    if "accelerate" not in sys.modules:
        class Accelerator:   # single-process stand-in (train_rqvae.py:67-72,153,182,195,212)
            def __init__(self, split_batches=True, mixed_precision="no"):
                import torch
                self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
                self.is_main_process = True
            def prepare(self, *objs):
                out = tuple(o.to(self.device) if hasattr(o, "parameters") else o for o in objs)
                return out if len(out) > 1 else out[0]
            def autocast(self):
                import contextlib
                return contextlib.nullcontext()
            def backward(self, loss):
                loss.backward()
            def wait_for_everyone(self):
                pass
            def clip_grad_norm_(self, params, max_norm):
                import torch
                return torch.nn.utils.clip_grad_norm_(params, max_norm)
        _stub("accelerate", Accelerator=Accelerator)
    if "torch_geometric" not in sys.modules:
        class _Base:
            def __init__(self, *a, **k):
                pass
        tg = _stub("torch_geometric")
        tg.data = _stub("torch_geometric.data", InMemoryDataset=_Base, HeteroData=dict,
                        download_google_url=None, download_url=None, extract_zip=None)
        tg.io = _stub("torch_geometric.io", fs=None)
        tg.datasets = _stub("torch_geometric.datasets", MovieLens1M=_Base)
    if "polars" not in sys.modules:
        _stub("polars")
    if "sentence_transformers" not in sys.modules:
        _stub("sentence_transformers", SentenceTransformer=type("SentenceTransformer", (), {}))


def load():
    """Import the reference hot-path modules; returns a namespace of them."""
    assert available(), f"reference not found at {REFERENCE}"
    install_stubs()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import importlib
    ns = types.SimpleNamespace()
    ns.quantize = importlib.import_module("modules.quantize")
    ns.rqvae = importlib.import_module("modules.rqvae")
    ns.kmeans = importlib.import_module("init.kmeans")
    ns.gumbel = importlib.import_module("distributions.gumbel")
    ns.encoder = importlib.import_module("modules.encoder")
    ns.loss = importlib.import_module("modules.loss")
    ns.normalize = importlib.import_module("modules.normalize")
    ns.schemas = importlib.import_module("data.schemas")
    ns.semids = importlib.import_module("modules.tokenizer.semids")
    return ns
