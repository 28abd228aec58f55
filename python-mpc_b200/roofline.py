"""Algorithmic HBM bytes of the kernels — used by bench.py for the roofline."""


def admm_elements_per_stage(nx, nu, slack):
    """Elements one thread reads + writes per stage and per ADMM iteration in admm_one (csrc/qp_thread.cuh).

    forward sweep : scaling vectors of the stage (D_x, E_bx, E_dyn(k+1), D_s, D_u, E_bu), iterates x_k/s_k/u_k,
                    (z, y) of rows bx_k, bu_k, dyn_{k+1}, the factor blocks Linv_k and F_{k-1}; writes t_k
    backward sweep: the same scaling vectors + E_dyn(k) + D_x(k+1), t_k, Linv_k and F_k, old x/z/y; writes new x/z/y
    """
    ns = nx if slack else 0
    nw, vs, cs = nx + nu, nx + ns + nu, 2 * nx + nu
    fac = nw * (nw + 1) // 2 + nx * nw
    coef = 3 * nx + ns + 2 * nu
    reads = 2 * coef + 2 * nx + 2 * vs + 4 * cs + 2 * fac + nw
    writes = nw + vs + 2 * cs
    return reads, writes


def admm_bytes_per_qp_iteration(N, nx, nu, slack, elem_size):
    r, w = admm_elements_per_stage(nx, nu, slack)
    return (N + 1) * (r + w) * elem_size
