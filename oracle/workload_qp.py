"""ORACLE — test infrastructure only.  Builds, on the CPU and in float64, the QP that the product's
synthetic workloads (python-mpc_b200/workloads.py) describe, for checking the GPU path.

The lateral bicycle discretisation is restated with scipy.linalg.expm (the zero-order-hold method of
the reference's own commented discretisation, Vehicle_Dynamics/vehicle_models.py:296-315 and
combined_longitudinal_lateral_dynamics.py:264-281)."""
import numpy as np
from scipy.linalg import expm

from . import ref_qp


def lateral_models(v, m=1300., l_f=1.25, l_r=1.40, Iz=2555.88174, Cf=11979.9261, Cr=11140.9949, dt=0.02):
    """lateral_model for an array of speeds (one batched scipy expm) -> (B,4,4), (B,4,1)."""
    v = np.asarray(v, dtype=np.float64).copy()
    vmin = 0.1
    v[(v >= 0) & (v < vmin)] = vmin
    v[(v < 0) & (v > -vmin)] = -vmin
    M = np.zeros((v.size, 5, 5))
    M[:, 0, 0] = -(Cf + Cr) / (m * v); M[:, 0, 1] = (Cr * l_r - Cf * l_f) / (m * v * v) - 1.0
    M[:, 1, 0] = (Cr * l_r - Cf * l_f) / Iz; M[:, 1, 1] = -(Cf * l_f ** 2 + Cr * l_r ** 2) / (Iz * v)
    M[:, 2, 1] = 1.0; M[:, 3, 0] = v; M[:, 3, 2] = v
    M[:, 0, 4] = Cf / (m * v); M[:, 1, 4] = Cf * l_f / Iz
    E = expm(M * dt)
    return E[:, :4, :4], E[:, :4, 4:5]


def lateral_model(v, m=1300., l_f=1.25, l_r=1.40, Iz=2555.88174, Cf=11979.9261, Cr=11140.9949, dt=0.02):
    vmin = 0.1
    if 0 <= v < vmin:
        v = vmin
    if -vmin < v < 0:
        v = -vmin
    A = np.array([[-(Cf + Cr) / (m * v), (Cr * l_r - Cf * l_f) / (m * v * v) - 1.0, 0, 0],
                  [(Cr * l_r - Cf * l_f) / Iz, -(Cf * l_f ** 2 + Cr * l_r ** 2) / (Iz * v), 0, 0],
                  [0, 1.0, 0, 0],
                  [v, 0, v, 0]])
    B = np.array([Cf / (m * v), Cf * l_f / Iz, 0, 0])
    M = np.zeros((5, 5)); M[:4, :4] = A; M[:4, 4] = B
    E = expm(M * dt)
    return E[:4, :4], E[:4, 4:5]


def lateral_qp(wl, b):
    """MpcQp of QP b of a workloads.LateralWorkload."""
    Ad, Bd = lateral_model(float(wl.speed[b]))
    if wl.increment:
        At, Bt, _ = ref_qp.augment_increment(Ad, Bd, None)
        Q = np.concatenate([wl.Q, [0.0]]); xr = np.concatenate([wl.xr[b], [0.0]])
    else:
        At, Bt, Q, xr = Ad, Bd, wl.Q, wl.xr[b]
    return ref_qp.canonical(wl.N, At, Bt, None, Q, Q, wl.R, xr, wl.xmin, wl.xmax, wl.umin, wl.umax, wl.x0[b],
                            slack=wl.slack, W=wl.W if wl.slack else None, S=wl.S if wl.slack else None)


def stage_perm(p):
    """Fill-reducing ordering for oracle/osqp_admm.c: variables interleaved by stage."""
    perm = np.zeros(p.nvar, dtype=np.int32)
    pos = 0
    for k in range(p.N + 1):
        for i in range(p.nx):
            perm[p.ix(k, i)] = pos; pos += 1
        if p.slack:
            for i in range(p.nx):
                perm[p.is_(k, i)] = pos; pos += 1
        if k < p.N:
            for i in range(p.nu):
                perm[p.iu(k, i)] = pos; pos += 1
    return perm


def lateral_batch_csc(wl, count=None):
    """All QPs of a workload as one shared CSC pattern + per-QP value rows (input of oracle_solve_batch).
    Vectorised: the positions of the model entries inside the CSC value array are found once, by assembling
    QP 0 with tagged model entries."""
    import scipy.sparse as sp
    B = wl.B if count is None else min(count, wl.B)
    p0 = lateral_qp(wl, 0)
    N, nx, nu = p0.N, p0.nx, p0.nu
    tagA = 1000.0 + np.arange(N * nx * nx, dtype=np.float64).reshape(N, nx, nx)
    tagB = 1e6 + np.arange(N * nx * nu, dtype=np.float64).reshape(N, nx, nu)
    pt = ref_qp.canonical(N, tagA, tagB, None, p0.Q, p0.QN, p0.R, p0.Xr, p0.xmin, p0.xmax, p0.umin, p0.umax, p0.x_init,
                          slack=p0.slack, W=p0.W, S=p0.S)
    P, q0, At, l0, u0 = ref_qp.assemble(pt)
    At = sp.csc_matrix(At); At.sort_indices()
    Pu = sp.triu(sp.csc_matrix(P), format="csc"); Pu.sort_indices()
    data = At.data
    posA = np.zeros((N, nx, nx), dtype=np.int64); posB = np.zeros((N, nx, nu), dtype=np.int64)
    isA = (data >= 1000.0) & (data < 1e6); isB = data >= 1e6
    posA.ravel()[(data[isA] - 1000.0).astype(np.int64)] = np.nonzero(isA)[0]
    posB.ravel()[(data[isB] - 1e6).astype(np.int64)] = np.nonzero(isB)[0]
    Ad, Bd = lateral_models(wl.speed[:B])
    if wl.increment:
        Ad, Bd, _ = ref_qp.augment_increment(Ad, Bd, None)
    Av = np.tile(np.where(isA | isB, 0.0, data), (B, 1))
    Av[:, posA.ravel()] = np.tile(Ad.reshape(B, 1, nx * nx), (1, N, 1)).reshape(B, -1)
    Av[:, posB.ravel()] = np.tile(Bd.reshape(B, 1, nx * nu), (1, N, 1)).reshape(B, -1)
    Pv = np.tile(Pu.data, (B, 1))
    q = np.tile(q0, (B, 1))                       # xr = 0 in the synthetic workloads -> q identical
    if np.any(wl.xr != 0):
        for b in range(B):
            q[b] = ref_qp.assemble(lateral_qp(wl, b))[1]
    l = np.tile(l0, (B, 1)); u = np.tile(u0, (B, 1))
    l[:, :nx] = -wl.x0[:B]; u[:, :nx] = -wl.x0[:B]
    At.data[:] = 1.0
    return Pu, At, Pv, q, Av, l, u, stage_perm(p0)


def dynamic_batch_csc(wl, A, Bm, g, Xr, idx):
    """QPs `idx` of a workloads.DynamicWorkload (configs[3]: per-stage linearisations A (B,N,6,6), Bm (B,N,6,2), g (B,N,6)
    — the matrices the device rollout produced, brought to the host — and stage references Xr (B,6,N+1)) as one shared
    CSC pattern + per-QP value rows (input of oracle_solve_batch)."""
    import scipy.sparse as sp
    idx = np.asarray(idx)
    nb = idx.size
    N, nx, nu = wl.N, 6, 2
    tagA = 1000.0 + np.arange(N * nx * nx, dtype=np.float64).reshape(N, nx, nx)
    tagB = 1e6 + np.arange(N * nx * nu, dtype=np.float64).reshape(N, nx, nu)
    p0 = ref_qp.canonical(N, tagA, tagB, g[idx[0]], wl.Q, wl.QN, wl.R, Xr[idx[0]], wl.xmin, wl.xmax, wl.umin, wl.umax,
                          wl.x0[idx[0]])
    P, q0, At, l0, u0 = ref_qp.assemble(p0)
    At = sp.csc_matrix(At); At.sort_indices()
    Pu = sp.triu(sp.csc_matrix(P), format="csc"); Pu.sort_indices()
    data = At.data
    isA = (data >= 1000.0) & (data < 1e6); isB = data >= 1e6
    posA = np.zeros(N * nx * nx, dtype=np.int64); posB = np.zeros(N * nx * nu, dtype=np.int64)
    posA[(data[isA] - 1000.0).astype(np.int64)] = np.nonzero(isA)[0]
    posB[(data[isB] - 1e6).astype(np.int64)] = np.nonzero(isB)[0]
    Av = np.tile(np.where(isA | isB, 0.0, data), (nb, 1))
    Av[:, posA] = A[idx].reshape(nb, -1)
    Av[:, posB] = Bm[idx].reshape(nb, -1)
    Pv = np.tile(Pu.data, (nb, 1))
    q = np.zeros((nb, p0.nvar)); l = np.tile(l0, (nb, 1)); u = np.tile(u0, (nb, 1))
    for k in range(N + 1):
        Qk = wl.QN if k == N else wl.Q
        q[:, k * nx:(k + 1) * nx] = -Qk[None, :] * Xr[idx][:, :, k]
    l[:, :nx] = -wl.x0[idx]; u[:, :nx] = -wl.x0[idx]
    for k in range(N):
        l[:, (k + 1) * nx:(k + 2) * nx] = -g[idx][:, k]; u[:, (k + 1) * nx:(k + 2) * nx] = -g[idx][:, k]
    At.data[:] = 1.0
    return Pu, At, Pv, q, Av, l, u, stage_perm(p0)
