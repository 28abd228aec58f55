"""Algorithmic HBM bytes of the kernels — used by bench.py for the roofline."""


def admm_elements_per_stage(nx, nu, slack):
    """Elements one thread reads / writes per stage and per ADMM iteration in admm_one (csrc/qp_thread.cuh, v2).

    forward sweep : scalings of the stage (D_x, D_s, D_u, E_bx, E_bu, E_dyn(k+1)), iterates x_k/s_k/u_k,
                    row state p of bx_k, bu_k, dyn_{k+1}, the factor block Linv_k;        writes t_k
    backward sweep: scalings (E_dyn(k) instead of E_dyn(k+1)), t_k, Linv_k, old x/s/u, old p;  writes new x/s/u, p
    """
    ns = nx if slack else 0
    nw, vs, cs = nx + nu, nx + ns + nu, 2 * nx + nu
    fac = nw * (nw + 1) // 2
    coef = 3 * nx + ns + 2 * nu
    reads = 2 * coef + 2 * vs + 2 * cs + 2 * fac + nw
    writes = nw + vs + cs
    return reads, writes


def admm_bytes_per_qp_iteration(N, nx, nu, slack, elem_size):
    r, w = admm_elements_per_stage(nx, nu, slack)
    return (N + 1) * (r + w) * elem_size
