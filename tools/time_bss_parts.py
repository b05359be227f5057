import time, torch
torch.manual_seed(0)
B, n = 32, 1024
A = torch.randn(B, n, n, dtype=torch.float64, device="cuda"); G = A @ A.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device="cuda")
D = torch.randn(B, n, 2, dtype=torch.float64, device="cuda")
def t(f, k=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
ref = torch.linalg.solve(G, D)
L = torch.linalg.cholesky_ex(G)[0]
def tri():
    y = torch.linalg.solve_triangular(L, D, upper=False)
    return torch.linalg.solve_triangular(L.transpose(1, 2), y, upper=True)
print("cholesky only ms", t(lambda: torch.linalg.cholesky_ex(G)[0]))
print("2x solve_triangular ms", t(tri), float((tri() - ref).abs().max()))
print("cholesky_solve ms", t(lambda: torch.cholesky_solve(D, L)))
print("cholesky_inverse+bmm ms", t(lambda: torch.cholesky_inverse(L) @ D), float((torch.cholesky_inverse(L) @ D - ref).abs().max()))
Dw = torch.randn(B, n, 64, dtype=torch.float64, device="cuda")
print("2x solve_triangular 64 rhs ms", t(lambda: torch.linalg.solve_triangular(L.transpose(1, 2), torch.linalg.solve_triangular(L, Dw, upper=False), upper=True)))
# 512 systems too
G5 = G[:, :512, :512].contiguous(); D5 = D[:, :512, :1].contiguous()
print("512: solve ms", t(lambda: torch.linalg.solve(G5, D5)))
L5 = torch.linalg.cholesky_ex(G5)[0]
print("512: chol ms", t(lambda: torch.linalg.cholesky_ex(G5)[0]), "tri ms", t(lambda: torch.linalg.solve_triangular(L5.transpose(1, 2), torch.linalg.solve_triangular(L5, D5, upper=False), upper=True)))
