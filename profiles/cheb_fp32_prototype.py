"""Numerics prototype (CPU, numpy/scipy on the oracle's assembled elasticity matrix) for the opt-in PE_CHEB_FP32=1 path:
Chebyshev(k)-Jacobi preconditioned CG where the passes INSIDE the polynomial read an FP32 copy of the matrix values and CG
itself stays FP64.  Prints CG iterations and matrix-pass equivalents (FP64 block-CSR pass = 8.44 B per scalar nonzero, FP32
pass = 4.44 B) against Jacobi-CG, from a zero start to the reference's absolute tolerance 1e-12 (DS:298).

    python profiles/cheb_fp32_prototype.py <refine>        # 3D Q1, 2^refine cells per axis

Output of 2026-10-18 (this container), see profiles/cheb_fp32_prototype.txt."""
import sys, time; sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent / 'tests'))
import helpers as H, numpy as np, scipy.sparse as sp
capi, fss = H.capi, H.fss
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H.load_oracle().po_set_threads(8)
inp = capi.InputData(text=H.make_input(dim=3, refine=refine, degree_u=1))
mesh = fss.make_mesh(inp)
b = H.create_oracle_backend()
dp, du, _ = fss.upload_problem(b, inp, mesh)
b.pressure_set_uniform(inp.p_init); b.displacement_assemble()
A = b.get_matrix(capi.MAT_ELASTICITY).tocsr(); rhs = b.get_vector(capi.VEC_U_RHS)
n = A.shape[0]; print("n", n, "nnz", A.nnz)
D = A.diagonal(); Dinv = 1.0/D
A32 = sp.csr_matrix((A.data.astype(np.float32).astype(np.float64), A.indices, A.indptr), shape=A.shape)
# lambda_max(D^-1 A) by power iteration (as the device does: 30 its x 1.1)
v = 0.5 + ((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % (1<<32) >> 22).astype(float)/1024.0
for _ in range(30):
    y = Dinv*(A@v); lam = np.linalg.norm(y)/np.linalg.norm(v); v = y/np.linalg.norm(y)
lam *= 1.1
def cheb(Aop, r, k, ratio=30.0):
    lmin = lam/ratio; theta = 0.5*(lam+lmin); delta = 0.5*(lam-lmin)
    sigma = theta/delta; rho = 1.0/sigma
    res = r.copy(); d = Dinv*res/theta; z = d.copy()
    for _ in range(k-1):
        rho_new = 1.0/(2*sigma - rho)
        res = res - Aop@d
        d = rho_new*rho*d + 2*rho_new/delta*(Dinv*res)
        z += d; rho = rho_new
    return z
def pcg(prec, tol=1e-12, maxit=5000):
    x = np.zeros(n); g = -rhs.copy(); it = 0
    h = prec(g); d = -h; gh = g@h
    while True:
        it += 1
        h = A@d; alpha = gh/(d@h); g += alpha*h; x += alpha*d
        if np.linalg.norm(g) <= tol or it >= maxit: break
        h = prec(g); beta = gh; gh = g@h; beta = gh/beta; d = beta*d - h
    return x, it
t=time.time(); xj, itj = pcg(lambda g: Dinv*g); print("Jacobi its", itj, "passes", itj, "%.1fs"%(time.time()-t))
for k in (2,3,4,6):
    x64, i64 = pcg(lambda g: cheb(A, g, k))
    x32, i32 = pcg(lambda g: cheb(A32, g, k))
    # cost model in FP64-BSR-pass equivalents: FP64 pass 8.44 B/nnz, FP32 pass 4.44 B/nnz ; vector ops per CG iteration ~0.12 of a pass, per cheb step ~0.06
    c64 = i64*(1 + (k-1)) ; c32 = i32*(1 + (k-1)*4.44/8.44)
    print(f"Cheb({k}): its fp64 {i64} fp32-matrix {i32}; matrix-pass equivalents {c64:.0f} vs {c32:.0f} (Jacobi {itj}); "
          f"|x32-xj|/|xj| {np.linalg.norm(x32-xj)/np.linalg.norm(xj):.2e}")
