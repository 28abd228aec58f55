"""ORACLE — test infrastructure only.  numpy restatement of Vehicle_Dynamics.get_dynamics_model
(/root/reference/Vehicle_Dynamics/vehicle_models.py:52-340): Pacejka lateral tyres, analytic Jacobians,
forward-Euler discretisation.  Checked against the reference itself through tests/golden/vehicle_models.npz
(tests/test_oracle.py)."""
import math

import numpy as np


def dynamics_model(x, u, m=1300., l_f=1.25, l_r=1.40, width=1.78, length=4.25, C_d=0.34, A_f=2.0, C_roll=0.015, dt=0.02):
    x = np.array(x, dtype=np.float64).ravel().copy(); u = np.array(u, dtype=np.float64).ravel().copy()
    wb = l_f + l_r
    Iz = 1 / 12 * m * (width ** 2 + length ** 2)
    roh = 1.23
    a = [-22.1, 1011, 1078, 1.82, 0.208]
    C = 1.30

    def tyre(Fz):
        D = a[0] * Fz ** 2 + a[1] * Fz
        BCD = a[2] * math.sin(a[3] * math.atan(a[4] * Fz))
        return D, BCD / (C * D) * 180 / np.pi

    Df, Bf = tyre(9.81 * (m * l_r / wb) * 0.001)
    Dr, Br = tyre(9.81 * (m * l_f / wb) * 0.001)
    if 0 <= x[3] < 0.5:                                   # vehicle_models.py:143-150
        x[4] = 0.; x[5] = 0.; u[0] = 0.
        if x[3] < 0.3:
            x[3] = 0.3
    if -0.5 < x[3] < 0:                                   # :152-159
        x[4] = 0.; x[5] = 0.; u[0] = 0.
        if x[3] > -0.3:
            x[3] = -0.3
    yaw, vx, vy, wz = x[2], x[3], x[4], x[5]
    st, acc = u
    af = -math.atan2(l_f * wz + vy, vx) + st
    ar = -math.atan2(-l_r * wz + vy, vx)
    Fyf = Df * math.sin(C * math.atan(Bf * af)); Fyr = Dr * math.sin(C * math.atan(Br * ar))
    sg = np.sign(vx)
    Fxf = m * acc - 0.5 * roh * C_d * A_f * vx ** 2 * sg - C_roll * m * 9.81 * sg
    sy, cy, ss, cs = math.sin(yaw), math.cos(yaw), math.sin(st), math.cos(st)
    f = np.array([vx * cy - vy * sy, vy * cy + vx * sy, wz,
                  1. / m * (Fxf * cs - Fyf * ss + m * vy * wz),
                  1. / m * (Fxf * ss + Fyr + Fyf * cs - m * vx * wz),
                  1. / Iz * (Fxf * l_f * ss + Fyf * l_f * cs - Fyr * l_r)])
    dFx = -roh * C_d * A_f * vx
    kf = (Bf * C * Df * math.cos(C * math.atan(Bf * af))) / (1 + Bf ** 2 * af ** 2)
    kr = (Br * C * Dr * math.cos(C * math.atan(Br * ar))) / (1 + Br ** 2 * ar ** 2)
    nf, nr = l_f * wz + vy, -l_r * wz + vy
    df, dr = nf ** 2 + vx ** 2, nr ** 2 + vx ** 2
    dFyf = np.array([kf * nf / df, kf * (-vx / df), kf * (-l_f * vx) / df]); dFyf_ds = kf
    dFyr = np.array([kr * nr / dr, kr * (-vx) / dr, kr * (l_r * vx) / dr])
    Ac = np.zeros((6, 6)); Bc = np.zeros((6, 2))
    Ac[0, 2:5] = [-vx * sy - vy * cy, cy, -sy]
    Ac[1, 2:5] = [-vy * sy + vx * cy, sy, cy]
    Ac[2, 5] = 1.
    Ac[3, 3:] = [1 / m * (dFx * cs - dFyf[0] * ss), 1 / m * (-dFyf[1] * ss + m * wz), 1 / m * (-dFyf[2] * ss + m * vy)]
    Ac[4, 3:] = [1 / m * (dFx * ss + dFyr[0] + dFyf[0] * cs - m * wz), 1 / m * (dFyr[1] + dFyf[1] * cs),
                 1 / m * (dFyr[2] + dFyf[2] * cs - m * vx)]
    Ac[5, 3:] = [1 / Iz * (dFx * l_f * ss + dFyf[0] * l_f * cs - dFyr[0] * l_r), 1 / Iz * (dFyf[1] * l_f * cs - dFyr[1] * l_r),
                 1 / Iz * (dFyf[2] * l_f * cs - dFyr[2] * l_r)]
    Bc[3] = [1 / m * (-Fxf * ss - dFyf_ds * ss - Fyf * cs), 1 / m * (m * cs)]
    Bc[4] = [1 / m * (Fxf * cs + dFyf_ds * cs - Fyf * ss), 1 / m * (m * ss)]
    Bc[5] = [1 / Iz * (Fxf * l_f * cs + dFyf_ds * l_f * cs - Fyf * l_f * ss), 1 / Iz * (m * l_f * ss)]
    gc = f - Ac @ x - Bc @ u
    return np.eye(6) + Ac * dt, Bc * dt, gc * dt
