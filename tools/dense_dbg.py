import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qpp-fusion-rag_b200"))
import torch
from qppfusion import ops
Q, N, D, k = 3452, 2_681_468, 768, 1000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
corpus = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
for s in range(0, N, 262144):
    x = torch.randn((min(262144, N - s), D), generator=g, device=dev)
    corpus[s:s + x.shape[0]] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
x = torch.randn((Q, D), generator=g, device=dev)
queries = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    sc, ids, st = torch.ops.qppfusion.dense_topk(queries, corpus, k, 0)
    torch.cuda.synchronize(); print("raw op", time.time() - t, "bad:", int((st != 0).sum()), st[st != 0][:10].tolist())
    t = time.time()
    sc, ids, st = ops.dense_topk(queries, corpus, k, 0)
    torch.cuda.synchronize(); print("wrapped", time.time() - t)
