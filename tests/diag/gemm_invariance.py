"""Is the fp32 cuBLAS GEMM of the trunk row-wise identical for M = B and M = B * W rows?  (decides whether teacher-forced
tiles can reproduce the step-wise logits bit for bit)"""
import torch
torch.manual_seed(0)
B, W = 1024, 32
for (K, N) in ((768, 2304), (768, 768), (768, 3072), (3072, 768), (768, 42001)):
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(B * W, K, device="cuda", dtype=dtype)
        w = torch.randn(K, N, device="cuda", dtype=dtype) * 0.02
        big = x @ w
        same = all(torch.equal(big[j * B:(j + 1) * B], x[j * B:(j + 1) * B] @ w) for j in range(0, W, 7))
        one = torch.equal(big[:1], x[:1] @ w), torch.equal(big[:8], x[:8] @ w)
        print(K, N, dtype, "M=%d slices == M=%d GEMM: %s ; M=1: %s, M=8: %s" % (B, B * W, same, one[0], one[1]))
x = torch.randn(B * W, 768, device="cuda")
g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
ln_big = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-5)
print("layer_norm row-invariant:", torch.equal(ln_big[:B], torch.nn.functional.layer_norm(x[:B], (768,), g, b, 1e-5)))
# attention: one query against the cache, alone vs inside a tile of W queries
H, hd, T = 12, 64, 256
q = torch.randn(B, H, W, hd, device="cuda"); k = torch.randn(B, H, T, hd, device="cuda"); v = torch.randn(B, H, T, hd, device="cuda")
s_big = q @ k.transpose(-1, -2)
s_one = q[:, :, 5:6] @ k.transpose(-1, -2)
print("q.K^T one query vs tile:", torch.equal(s_big[:, :, 5:6], s_one))
a = s_big.softmax(-1)
print("att.V one query vs tile:", torch.equal((a @ v)[:, :, 5:6], a[:, :, 5:6] @ v))
