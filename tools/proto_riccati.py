"""Developer prototype (numpy) of the stage-wise Riccati sweep the CUDA solver implements.
Checks that one sweep reproduces the dense KKT Newton step of oracle/nlp.py.  Not shipped,
not imported by the product; run by hand:  python tools/proto_riccati.py
"""
import numpy as np
from oracle import nlp


def riccati_step(P, V, p, lam, delta=0.0):
    nx, nu, N, s, h = P.nx, P.nu, P.N, P.s, P.h
    X, U = P.split_V(V)
    pp = P.split_p(p)
    Q, R, Rm, r, u_init = pp["Q"], pp["R"], pp["Rm"], pp["r"], pp["u_init"]
    F, Fx, Fu = P.step_jac(X[:-1], U, pp)
    e = F - r
    lamk = lam.reshape(N, nx)
    W = 2 * Q * e + lamk
    Hc = P.step_hess(X[:-1], U, pp, W)
    Uprev = np.vstack([u_init[None], U[:-1]])
    dU = U - Uprev
    Pxx = np.zeros((nx, nx)); Pxv = np.zeros((nx, nu)); Pvv = np.zeros((nu, nu))
    px = np.zeros(nx); pv = np.zeros(nu)
    Ks = [None] * N
    Ps = [None] * (N + 1)
    Ps[N] = (Pxx.copy(), Pxv.copy(), px.copy())
    R2 = np.diag(2 * R)
    for k in range(N - 1, -1, -1):
        A, B = Fx[k], Fu[k]
        c = F[k] - X[k + 1]
        Fz = np.hstack([A, B])
        Hzz = 2 * (Fz.T * Q) @ Fz + Hc[k]
        Hzz[nx:, nx:] += np.diag(2 * Rm) + R2 + delta * np.eye(nu)
        Hzz[:nx, :nx] += delta * np.eye(nx)
        qe = 2 * Q * e[k]
        qx = A.T @ qe
        qu = B.T @ qe + 2 * R * dU[k] + 2 * Rm * U[k]
        qv = -2 * R * dU[k]
        y = Pxx @ c + px
        yv = Pxv.T @ c + pv
        PB = Pxx @ B + Pxv
        Gxx = Hzz[:nx, :nx] + A.T @ Pxx @ A
        Gxu = Hzz[:nx, nx:] + A.T @ PB
        Guu = Hzz[nx:, nx:] + B.T @ PB + Pxv.T @ B + Pvv
        gx = qx + A.T @ y
        gu = qu + B.T @ y + yv
        Gvu = -R2
        L = np.linalg.cholesky(Guu)
        Kx = -np.linalg.solve(Guu, Gxu.T)
        Kv = -np.linalg.solve(Guu, Gvu.T)
        k0 = -np.linalg.solve(Guu, gu)
        Ks[k] = (Kx, Kv, k0)
        Pxx = Gxx + Gxu @ Kx
        Pxx = 0.5 * (Pxx + Pxx.T)
        Pxv = Gxu @ Kv
        Pvv = R2 + Gvu @ Kv
        px = gx + Gxu @ k0
        pv = qv + Gvu @ k0
        Ps[k] = (Pxx.copy(), Pxv.copy(), px.copy())
    dx = np.zeros(nx); dv = np.zeros(nu)
    dX = np.zeros((N + 1, nx)); dUU = np.zeros((N, nu)); lam_new = np.zeros((N, nx))
    for k in range(N):
        Kx, Kv, k0 = Ks[k]
        du = Kx @ dx + Kv @ dv + k0
        dxn = Fx[k] @ dx + Fu[k] @ du + (F[k] - X[k + 1])
        dUU[k] = du
        dX[k + 1] = dxn
        Pxx, Pxv, px = Ps[k + 1]
        lam_new[k] = Pxx @ dxn + Pxv @ du + px
        dx, dv = dxn, du
    return P.join_V(dX, dUU), lam_new.ravel()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    P = nlp.Problem(N=12)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, P.N, P.h, 1)
    p = P.pack_p(traj[0], Q, R, Rm, np.array([0.3, -0.2]))
    V = rng.uniform(-0.3, 0.3, P.NV); V[:4] = x0[0]
    lam = rng.uniform(-1, 1, P.M)
    d, ln = riccati_step(P, V, p, lam)
    # dense KKT
    f, gf = P.nlp_grad_f(V, p); g, J = P.nlp_jac_g(V, p); H = P.nlp_hess_l(V, p, 1.0, lam)
    fr = np.arange(4, P.NV)
    K = np.block([[H[np.ix_(fr, fr)], J[:, fr].T], [J[:, fr], np.zeros((P.M, P.M))]])
    sol = np.linalg.solve(K, -np.concatenate([gf[fr], g]))
    print("step err", np.abs(sol[:fr.size] - d[4:]).max() / np.abs(d).max(),
          "lam err", np.abs(sol[fr.size:] - ln).max() / np.abs(ln).max())
