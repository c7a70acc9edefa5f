"""BASELINE config[1]/[2]-shaped pipeline: items x 768 -> encoder MLP (768-512-256-128-32) -> 3-level RQ at D=32 -> ids.
Codebooks are k-means-initialised on the encoder outputs (live codes), so the fp32-vs-bf16 id agreement is meaningful."""
import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200.modules.rqvae import RqVae
from rq_vae_recommender_b200.modules.quantize import QuantizeForwardMode
from rq_vae_recommender_b200.data.schemas import SeqBatch
torch.manual_seed(0); np.random.seed(0)
m = RqVae(input_dim=768, embed_dim=32, hidden_dims=[512, 256, 128], codebook_size=256, codebook_kmeans_init=True,
          codebook_mode=QuantizeForwardMode.STE, n_layers=3, n_cat_features=0).cuda()
# clustered synthetic items so the codes are live (like a trained model), then the lazy k-means init (train_rqvae.py:178-183)
centers = I.unit_rows(40, 200, 768)
def items(n, seed):
    v = centers[np.random.RandomState(seed).randint(0, 200, n)] + 0.5 * I.unit_rows(seed + 1, n, 768)
    return torch.from_numpy((v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)).cuda()
m.train()
with torch.no_grad():
    m(SeqBatch(None, None, None, items(20000, 5), None, None), 0.2)
m.eval()
assert all(l.kmeans_initted for l in m.layers)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

FLOP = 2 * 561152
with torch.no_grad():                                   # the bf16 path is forward-only: everything below is no_grad
    for N in (12101, 65536):
        x = items(N, 100 + N)
        t32 = timeit(lambda: m.tokenize(x)); t16 = timeit(lambda: m.tokenize(x, mlp_precision="bf16"))
        m.encoder.precision = "fp32"; e32 = timeit(lambda: m.encode(x))
        m.encoder.precision = "bf16"; e16 = timeit(lambda: m.encode(x)); m.encoder.precision = "fp32"
        ids32, ids16 = m.tokenize(x), m.tokenize(x, mlp_precision="bf16")
        nuniq = torch.unique(ids32, dim=0).shape[0]
        per_level = [(ids32[:, l] == ids16[:, l]).float().mean().item() * 100 for l in range(3)]
        print(f"N={N}: fp32 tokenize {t32:.3f} ms ({N/t32/1e3:.1f} M items/s), encoder {e32:.3f} ms ({FLOP*N/e32/1e9:.0f} TFLOP/s) | "
              f"bf16 tokenize {t16:.3f} ms ({N/t16/1e3:.1f} M items/s), encoder {e16:.3f} ms ({FLOP*N/e16/1e9:.0f} TFLOP/s) | "
              f"unique id tuples {nuniq}; bf16 ids == fp32 ids: rows {(ids32==ids16).all(1).float().mean().item()*100:.1f}%, per level "
              + "/".join(f"{p:.1f}" for p in per_level) + "%")
