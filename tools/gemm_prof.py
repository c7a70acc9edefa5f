import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
x = torch.from_numpy(I.unit_rows(1, 65536, 768)).cuda()
ws = [torch.from_numpy(w).cuda() for w in I.mlp_weights(2, [768, 512, 256, 128, 32])]
wi = [ops.to_bf16_image(w) for w in ws]
for _ in range(3): ops.mlp_forward_bf16(x, ws, weight_images=wi)
torch.cuda.synchronize(); print("done")
