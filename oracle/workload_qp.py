"""ORACLE — test infrastructure only.  Builds, on the CPU and in float64, the QP that the product's
synthetic workloads (python-mpc_b200/workloads.py) describe, for checking the GPU path.

The lateral bicycle discretisation is restated with scipy.linalg.expm (the zero-order-hold method of
the reference's own commented discretisation, Vehicle_Dynamics/vehicle_models.py:296-315 and
combined_longitudinal_lateral_dynamics.py:264-281)."""
import numpy as np
from scipy.linalg import expm

from . import ref_qp


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
