import sys, torch
sys.path.insert(0, '/root/repo')
from gaudi_b200 import _lib, training
dev = torch.device('cuda:0')
def run(G, X, M, N):
    K = G.shape[0]
    C = torch.zeros(M, N, device=dev)
    _lib.check(_lib.lib().gb_wgrad(K, M, N, training._ptr(G), G.stride(0), training._ptr(X), X.stride(0), training._ptr(C), N, 0, training._stream()))
    torch.cuda.synchronize()
    return C
M = N = 64
K = 16
G = torch.zeros(K, M, device=dev); X = torch.zeros(K, N, device=dev)
G[0, 0] = 1.0; X[0] = torch.arange(1, N + 1, device=dev).float()
C = run(G, X, M, N)
print("nonzero count", int((C != 0).sum()), "C[0,:8]", C[0, :8].tolist())
nz = (C != 0).nonzero()
print(nz[:10].tolist(), C[C != 0][:10].tolist())
G = torch.randn(K, M, device=dev); X = torch.randn(K, N, device=dev)
C = run(G, X, M, N); ref = G.T @ X
print("rand max err", float((C - ref).abs().max()), "C norm", float(C.norm()), "ref norm", float(ref.norm()))
